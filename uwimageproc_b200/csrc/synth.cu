// synth.cu - deterministic integer-only synthetic underwater-like frames and frame checksums
// (SURVEY.md 8d).  CPU twin: oracle/uwip_oracle.py:synth_frame (tests compare them byte for byte).
#include <algorithm>

#include "common.cuh"

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ int tri_wave(int t, int period) {
  int p = t % period;
  int v = (p * 510) / period;
  return v > 255 ? 510 - v : v;
}
__device__ __forceinline__ int value_noise(uint32_t seed, uint32_t f, int x, int y, int cell, uint32_t salt) {
  int cx = x / cell, cy = y / cell;
  int fx = ((x % cell) * 256) / cell, fy = ((y % cell) * 256) / cell;
  uint32_t base = seed + salt * 0x9E3779B1u + f * 0x85EBCA77u;
  auto lat = [&](int ix, int iy) { return (int)(lowbias32(base + (uint32_t)ix * 0xC2B2AE3Du + (uint32_t)iy * 0x27D4EB2Fu) & 255u); };
  int h00 = lat(cx, cy), h10 = lat(cx + 1, cy), h01 = lat(cx, cy + 1), h11 = lat(cx + 1, cy + 1);
  int top = h00 * (256 - fx) + h10 * fx, bot = h01 * (256 - fx) + h11 * fx;
  return (top * (256 - fy) + bot * fy) >> 16;
}

// A thread writes four adjacent pixels (x0 a multiple of 4): they share the cells of both noise octaves (8 and 32 wide), so
// the eight lattice hashes and the row's depth term are formed once, and the twelve bytes leave as three words when the row
// pitch allows it.  Same integer arithmetic per pixel as value_noise / tri_wave above (the oracle twin is their definition).
__global__ void __launch_bounds__(256) synth_kernel(uint8_t* __restrict__ dst, uint32_t seed, int first, int W, int H) {
  const int f = first + blockIdx.z;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x0 >= W) return;
  const int px = max(W / 2, 2), py = max(H / 3, 2);
  const int wy = tri_wave(y + 5 * f, py);
  int p = (x0 + 7 * f) % px;
  // lattice corners of the two octaves
  int lat[2][4];
#pragma unroll
  for (int o = 0; o < 2; o++) {
    const int sh = o == 0 ? 5 : 3;   // cell = 32, 8
    const int cx = x0 >> sh, cy = y >> sh;
    const uint32_t base = seed + (uint32_t)(o + 1) * 0x9E3779B1u + (uint32_t)f * 0x85EBCA77u;
#pragma unroll
    for (int k = 0; k < 4; k++)
      lat[o][k] = (int)(lowbias32(base + (uint32_t)(cx + (k & 1)) * 0xC2B2AE3Du + (uint32_t)(cy + (k >> 1)) * 0x27D4EB2Fu) & 255u);
  }
  const int fy32 = (y & 31) * 8, fy8 = (y & 7) * 32;
  const int base[3] = {120, 140, 30}, gain[3] = {60, 50, -25}, texgain[3] = {40, 36, 16};
  const uint32_t krow = seed + (uint32_t)f * 0x85EBCA77u + (uint32_t)y * 0x27D4EB2Fu;
  uint8_t b[12];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int x = x0 + j;
    while (p >= px) p -= px;
    int v = (p * 510) / px;
    const int wx = v > 255 ? 510 - v : v;
    p++;
    const int depth = (wx + wy) >> 1;
    const int fx32 = (x & 31) * 8, fx8 = (x & 7) * 32;
    const int n32 = ((lat[0][0] * (256 - fx32) + lat[0][1] * fx32) * (256 - fy32) + (lat[0][2] * (256 - fx32) + lat[0][3] * fx32) * fy32) >> 16;
    const int n8 = ((lat[1][0] * (256 - fx8) + lat[1][1] * fx8) * (256 - fy8) + (lat[1][2] * (256 - fx8) + lat[1][3] * fx8) * fy8) >> 16;
    const int tex = 2 * n32 + n8 - 384;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint32_t k = krow + (uint32_t)x * 0xC2B2AE3Du + (uint32_t)(c + 1) * 0x165667B1u;
      const int noise = (int)(lowbias32(k) & 7u) - 3;
      const int val = base[c] + ((depth * gain[c]) >> 8) + ((tex * texgain[c]) >> 8) + noise;
      b[3 * j + c] = (uint8_t)min(max(val, 0), 255);
    }
  }
  uint8_t* o = dst + ((size_t)blockIdx.z * W * H + (size_t)y * W + x0) * 3;
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
    for (int k = 0; k < 3; k++) ow[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (x0 + j < W) { o[3 * j] = b[3 * j]; o[3 * j + 1] = b[3 * j + 1]; o[3 * j + 2] = b[3 * j + 2]; }
  }
}

int synth_frames_dev(uwip_ctx* ctx, uint8_t* d_dst, uint32_t seed, int first, int n, int w, int h) {
  dim3 grid(cdiv(w, 1024), h, n);   // four pixels per thread
  UWIP_LAUNCH(ctx, "synth", synth_kernel, grid, 256, 0, d_dst, seed, first, w, h);
  return UWIP_OK;
}

// checksum_f = sum_i (byte_i + 1) * ((i * 2654435761 mod 2^32) | 1)   (mod 2^64)
// sixteen bytes per load where the frame allows it (a frame of a batch starts wherever the one before ended: the bytes
// before the first 16-byte boundary and after the last one go one by one)
__device__ __forceinline__ unsigned long long checksum_byte(unsigned long long acc, uint32_t byte, uint32_t i) {
  return acc + (unsigned long long)(byte + 1u) * ((i * 2654435761u) | 1u);
}
__global__ void __launch_bounds__(256) checksum_kernel(const uint8_t* __restrict__ src, size_t n_bytes, unsigned long long* sums) {
  const uint8_t* p = src + (size_t)blockIdx.y * n_bytes;
  size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
  if (head > n_bytes) head = n_bytes;
  const size_t nvec = (n_bytes - head) / 16;
  const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nthr = (size_t)gridDim.x * 256;
  const uint4* pv = reinterpret_cast<const uint4*>(p + head);
  unsigned long long acc = 0;
  for (size_t v = tid; v < nvec; v += nthr) {
    const uint4 q = __ldg(pv + v);
    const uint32_t i0 = (uint32_t)(head + v * 16);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 16; j++) acc = checksum_byte(acc, (w[j >> 2] >> (8 * (j & 3))) & 255u, i0 + (uint32_t)j);
  }
  for (size_t i = tid; i < head; i += nthr) acc = checksum_byte(acc, p[i], (uint32_t)i);
  for (size_t i = head + nvec * 16 + tid; i < n_bytes; i += nthr) acc = checksum_byte(acc, p[i], (uint32_t)i);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sums[blockIdx.y], acc);
}

int checksum_frames_dev(uwip_ctx* ctx, const uint8_t* d_src, int n, int w, int h, uint64_t* sums_host) {
  unsigned long long* d_sums = (unsigned long long*)uwip_slot(ctx, SLOT_MISC, std::max<size_t>(256, (size_t)n * 8));
  if (!d_sums) return UWIP_ERR_NOMEM;
  UWIP_CUDA(ctx, cudaMemsetAsync(d_sums, 0, (size_t)n * 8, ctx->stream));
  size_t n_bytes = (size_t)w * h * 3;
  dim3 grid(std::max(1, std::min(1024, ctx->sm_count * 8 / n + 1)), n);
  UWIP_LAUNCH(ctx, "checksum", checksum_kernel, grid, 256, 0, d_src, n_bytes, d_sums);
  UWIP_CUDA(ctx, cudaMemcpyAsync(sums_host, d_sums, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  UWIP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UWIP_OK;
}

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

__global__ void __launch_bounds__(256) synth_kernel(uint8_t* __restrict__ dst, uint32_t seed, int first, int W, int H) {
  int f = first + blockIdx.z;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  int px = max(W / 2, 2), py = max(H / 3, 2);
  int depth = (tri_wave(x + 7 * f, px) + tri_wave(y + 5 * f, py)) >> 1;
  int tex = 2 * value_noise(seed, (uint32_t)f, x, y, 32, 1) + value_noise(seed, (uint32_t)f, x, y, 8, 2) - 384;
  const int base[3] = {120, 140, 30}, gain[3] = {60, 50, -25}, texgain[3] = {40, 36, 16};
  uint8_t* o = dst + ((size_t)blockIdx.z * W * H + (size_t)y * W + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    uint32_t k = seed + (uint32_t)f * 0x85EBCA77u + (uint32_t)y * 0x27D4EB2Fu + (uint32_t)x * 0xC2B2AE3Du + (uint32_t)(c + 1) * 0x165667B1u;
    int noise = (int)(lowbias32(k) & 7u) - 3;
    int v = base[c] + ((depth * gain[c]) >> 8) + ((tex * texgain[c]) >> 8) + noise;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

int synth_frames_dev(uwip_ctx* ctx, uint8_t* d_dst, uint32_t seed, int first, int n, int w, int h) {
  dim3 grid(cdiv(w, 256), h, n);
  UWIP_LAUNCH(ctx, "synth", synth_kernel, grid, 256, 0, d_dst, seed, first, w, h);
  return UWIP_OK;
}

// checksum_f = sum_i (byte_i + 1) * ((i * 2654435761 mod 2^32) | 1)   (mod 2^64)
__global__ void __launch_bounds__(256) checksum_kernel(const uint8_t* __restrict__ src, size_t n_bytes, unsigned long long* sums) {
  const uint8_t* p = src + (size_t)blockIdx.y * n_bytes;
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_bytes; i += (size_t)gridDim.x * 256) {
    uint32_t wgt = ((uint32_t)i * 2654435761u) | 1u;
    acc += (unsigned long long)(p[i] + 1u) * wgt;
  }
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

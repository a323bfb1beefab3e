// blurmetric.cu - calcBlur (frame-quality gate in front of the chain) for sm_100a.
//   reference: modules/videostrip/src/videostrip.cpp:170-184 (calcBlur) and :39-60 (calcBlurGPU):
//     cvtColor(BGR2GRAY) -> Laplacian(grey, laplacian, grey.type(), CV_16S) -> meanStdDev -> stdev.
//   The third argument of that Laplacian call is the output depth (grey.type() = CV_8U) and the fourth is the
//   APERTURE (CV_16S == 3): kernel [2 0 2; 0 -8 0; 2 0 2], reflect-101 border, result saturated to 8 bits.
//   Everything up to the two sums is integer work and bit-exact; mean / stdev follow cv::meanStdDev
//   (scale = 1/N; mean = s*scale; var = max(sq*scale - mean^2, 0)) in IEEE double without contraction.
// One pass over the frame (3 B/px): a warp walks down a strip of 30 columns (+1 halo lane on both sides),
// keeps the three-row window in registers and exchanges the horizontal neighbours by shuffle.
#include "common.cuh"

constexpr int LB_WARPS = 8, LB_COLS = 30, LB_ROWS = 64;

__device__ __forceinline__ int refl101(int i, int n) {
  if (n == 1) return 0;
  return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}
// cvtColor(BGR2GRAY) 8-bit of cv2 4.13.0: Q15 (equal on all 2^24 triples, tests/golden/kat.json: all_bgr2gray_crc)
__device__ __forceinline__ int bgr2gray_u8(int b, int g, int r) { return (9798 * r + 19235 * g + 3735 * b + 16384) >> 15; }

// grid (strips of 8 x 30 columns, bands of 64 rows, frames); sums[2f] += sum L, sums[2f+1] += sum L^2
// AP: aperture.  3 = calcBlur's kernel [2 0 2; 0 -8 0; 2 0 2]; 1 = [0 1 0; 1 -4 1; 0 1 0], the kernel
// calcBlurGPU asks cv::cuda::createLaplacianFilter for (videostrip.cpp:48).
template <int AP>
__global__ void __launch_bounds__(32 * LB_WARPS) lap_moments_kernel(const uint8_t* __restrict__ src, int W, int H,
                                                                    unsigned long long* __restrict__ sums,
                                                                    uint8_t* __restrict__ lap_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = (blockIdx.x * LB_WARPS + warp) * LB_COLS + lane - 1;  // lanes 0 and 31 are the halo columns
  const int y0 = blockIdx.y * LB_ROWS, y1 = min(y0 + LB_ROWS, H);
  const uint8_t* fr = src + (size_t)blockIdx.z * W * H * 3;
  const bool out_lane = lane >= 1 && lane <= LB_COLS && x < W;
  const int xr = refl101(min(x, W), W);  // x == -1 -> 1, x == W -> W - 2; lanes further right read a valid column and are ignored
  auto grey_at = [&](int y) {
    const uint8_t* p = fr + ((size_t)refl101(y, H) * W + xr) * 3;
    return bgr2gray_u8(p[0], p[1], p[2]);
  };
  auto diag = [&](int c) {  // g(x-1) + g(x+1) of the same row
    int l = __shfl_up_sync(0xffffffffu, c, 1), r = __shfl_down_sync(0xffffffffu, c, 1);
    return l + r;
  };
  uint32_t s1 = 0;
  unsigned long long s2 = 0;
  if ((blockIdx.x * LB_WARPS + warp) * LB_COLS < W) {  // warp-uniform
    int c_prev = grey_at(y0 - 1);
    int d_prev = diag(c_prev);
    int c_cur = grey_at(y0);
    int d_cur = diag(c_cur);
    for (int y = y0; y < y1; y++) {
      int c_next = grey_at(y + 1);
      int d_next = diag(c_next);
      int L = (AP == 3) ? 2 * (d_prev + d_next) - 8 * c_cur : (c_prev + c_next + d_cur) - 4 * c_cur;
      L = min(max(L, 0), 255);
      if (out_lane) {
        s1 += (uint32_t)L;
        s2 += (uint32_t)(L * L);
        if (lap_out) lap_out[((size_t)blockIdx.z * H + y) * W + x] = (uint8_t)L;
      }
      d_prev = d_cur; d_cur = d_next; c_prev = c_cur; c_cur = c_next;
    }
  }
  // block reduction: 64 rows x 255 < 2^32 per lane for s1
  unsigned long long a = s1, b = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  __shared__ unsigned long long sh[2][LB_WARPS];
  if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < LB_WARPS; k++) t += sh[threadIdx.x][k];
    if (t) atomicAdd(&sums[2 * blockIdx.z + threadIdx.x], t);
  }
}

// cv::meanStdDev on the 8-bit Laplacian: out[2f] = mean, out[2f+1] = stdev (double; calcBlur returns float(stdev))
__global__ void lap_finish_kernel(const unsigned long long* __restrict__ sums, int n, double n_px, double* __restrict__ out) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  double scale = __ddiv_rn(1.0, n_px);
  double mean = __dmul_rn((double)sums[2 * f], scale);
  double var = __dsub_rn(__dmul_rn((double)sums[2 * f + 1], scale), __dmul_rn(mean, mean));
  out[2 * f] = mean;
  out[2 * f + 1] = __dsqrt_rn(var > 0.0 ? var : 0.0);
}

int calc_blur_frames_dev(uwip_ctx* ctx, const uint8_t* d_src, int n, int w, int h, int aperture, double* d_mean_std, uint8_t* d_lap) {
  unsigned long long* d_sums = (unsigned long long*)uwip_slot(ctx, SLOT_BLURSUMS, (size_t)n * 16);
  if (!d_sums) return UWIP_ERR_NOMEM;
  UWIP_CUDA(ctx, cudaMemsetAsync(d_sums, 0, (size_t)n * 16, ctx->stream));
  dim3 grid(cdiv(w, LB_WARPS * LB_COLS), cdiv(h, LB_ROWS), n);
  if (aperture == 3) UWIP_LAUNCH(ctx, "lap_moments", lap_moments_kernel<3>, grid, 32 * LB_WARPS, 0, d_src, w, h, d_sums, d_lap);
  else UWIP_LAUNCH(ctx, "lap_moments", lap_moments_kernel<1>, grid, 32 * LB_WARPS, 0, d_src, w, h, d_sums, d_lap);
  UWIP_LAUNCH(ctx, "lap_finish", lap_finish_kernel, cdiv(n, 128), 128, 0, d_sums, n, (double)((size_t)w * h), d_mean_std);
  return UWIP_OK;
}

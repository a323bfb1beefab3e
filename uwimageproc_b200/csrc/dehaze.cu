// dehaze.cu - bgdehaze (Li et al. ICASSP-16 blue-green dehazing + red-channel correction + adaptive
// exposure map) for sm_100a.
//   reference: modules/bgdehaze/BGDehaze.py:14-89, modules/bgdehaze/guidedfilter.py:23-103,
//              modules/bgdehaze/main.py:16-19.  Stage names D0..D10 follow SURVEY.md 8a.
//
// Data flow per frame (all frame-global reductions land in FrameState, no host round trips):
//   minmax (D0)  ->  window: 15x15 window max / arg-min partials (D1), window min (D2), and the packed
//                    plane kq = (k'_b, k'_g, k'_r, m'_b) + plane m'_g that every later pass reads
//   -> GF1a: box sums of 17 moments + per-pixel 3x3 solve -> a,b planes (f32 x8)
//   -> GF1b: box(a,b) -> refined t (D3,D5) -> J (D6) + min/max/sum reductions -> J planes (f32 x2)
//   -> E: restored -> R8/I8 -> YCrCb joint min/max (D7,D8 first half); packed plane ycc = (Yi,Cri,Cbi,Yj)
//   -> S table (256x256 f64 per frame: the exposure ratio is a function of two bytes)
//   -> GF2a: S + 13 moments + solve -> a,b (f32 x4)   -> GF2b: box(a,b) -> refined S, exposure min/max
//   -> final: normalise, x255, rint, saturate -> bgr8 (D8 second half, D10)
//
// Box sums ("quad march").  A CTA owns a strip of image columns (halo of r columns on both sides) and
// walks down the rows.  Every thread owns FOUR adjacent columns: the vertical running sums of all
// moments stay in registers (add the entering row, subtract the leaving row).  For every output row a
// thread publishes the inclusive prefix over its own quad (one 16-byte shared store per moment) and
// the quad total; a few threads turn the quad totals into a prefix over the strip; the window sum of
// column 4t+c is then  G[t+r/4-1] - G[t-r/4-1] - quad[t-r/4][c-1] + quad[t+r/4][c]  (two 16-byte and
// two 4-byte shared loads per moment for four pixels).  Guide moments are exact 32-bit integers (the
// guide is k/range with k uint8); everything involving the filtered signal accumulates in fp64.  The
// per-pixel solve works in "k units" (guide not divided by range) with eps_k = eps*range^2 on the
// exact integer numerators N*S_ij - S_i*S_j.
#include <algorithm>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// frame state
// ------------------------------------------------------------------------------------------------
__global__ void fs_reset_kernel(FrameState* fs, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FrameState s;
  memset(&s, 0, sizeof(s));
  s.kmin = 255; s.kmax = 0;
  s.jmin_key[0] = s.jmin_key[1] = ~0ull;
  s.jmax_key[0] = s.jmax_key[1] = 0ull;
  s.rmin = 255; s.rmax = 0;
  s.yi_min = s.yj_min = 255; s.yi_max = s.yj_max = 0;
  s.omin_key = ~0ull; s.omax_key = 0ull;
  fs[i] = s;
}
int frame_state_reset(uwip_ctx* ctx, FrameState* fs, int n) {
  UWIP_LAUNCH(ctx, "fs_reset", fs_reset_kernel, cdiv(n, 128), 128, 0, fs, n);
  return UWIP_OK;
}
FrameState* frame_state_get(uwip_ctx* ctx, int n) { return (FrameState*)uwip_slot(ctx, SLOT_FSTATE, sizeof(FrameState) * (size_t)n); }

// ------------------------------------------------------------------------------------------------
// small fp64 helpers
// ------------------------------------------------------------------------------------------------
// exact uint32 -> double: one I2F on the conversion pipe (the 2^52 magic-add form costs three issue slots)
__device__ __forceinline__ double u2d(uint32_t u) { return __uint2double_rn(u); }
// 1/x to ~1 ulp (MUFU seed + two Newton steps).  Used where the result feeds continuous arithmetic
// only; every place whose result is truncated to a byte uses the IEEE division.
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Round to a multiple of 2^-G (round-half-even) by the add/subtract of 1.5*2^(52-G); valid for |x| < 2^(51-G).
// Every value that enters a running sum is put on such a grid: the vertical running sums (add the entering
// row, subtract the leaving row) are then EXACT in fp64, so the result of a frame does not depend on where a
// CTA's vertical segment starts - i.e. not on the batch size, the sub-batch split or the GPU count.
template <int G>
__device__ __forceinline__ double grid_round(double x) {
  const double M = 6755399441055744.0 / (double)(1ull << G);  // 1.5 * 2^(52-G)
  return (x + M) - M;
}

// cp.async (LDGSTS): global -> shared without a register in between
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem, const void* g, unsigned bytes, uint64_t* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(g), "r"(bytes), "r"(b) : "memory");
}

// ------------------------------------------------------------------------------------------------
// D0: joint min / max over all channels (bgdehaze/main.py:17)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) minmax_kernel(const uint8_t* __restrict__ src, size_t n_bytes, FrameState* fs) {
  const uint8_t* p = src + (size_t)blockIdx.y * n_bytes;
  unsigned int mn = 0x00ff00ffu, mx = 0;
  bool vec = ((((uintptr_t)p) & 15) == 0);
  size_t n16 = vec ? n_bytes / 16 : 0;
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < n16; g += (size_t)gridDim.x * 256) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + g);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      unsigned int e = w[k] & 0x00ff00ffu, o = (w[k] >> 8) & 0x00ff00ffu;
      mn = __vminu2(mn, __vminu2(e, o));
      mx = __vmaxu2(mx, __vmaxu2(e, o));
    }
  }
  unsigned int smn = min(mn & 0xffffu, mn >> 16), smx = max(mx & 0xffffu, mx >> 16);
  for (size_t i = n16 * 16 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n_bytes; i += (size_t)gridDim.x * 256) {
    smn = min(smn, (unsigned)p[i]);
    smx = max(smx, (unsigned)p[i]);
  }
  smn = warp_reduce_min_u32(smn);
  smx = warp_reduce_max_u32(smx);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&fs[blockIdx.y].kmin, smn);
    atomicMax(&fs[blockIdx.y].kmax, smx);
  }
}

// ------------------------------------------------------------------------------------------------
// D1 / D2: 15x15 window max (3 channels) -> arg-min partials; window min (blue, green); packed planes
// ------------------------------------------------------------------------------------------------
constexpr int WK_TX = 64, WK_TY = 32, WK_THREADS = 256;
constexpr int WK_MAXWIN = 33;
constexpr int WK_TWIN = 15;  // transmission window: always 15 (BGDehaze.py:52 drops w)

struct ArgPartial {
  double d0, d1;
  unsigned int i0, i1;
};

__device__ __forceinline__ bool lex_less(double a, unsigned ia, double b, unsigned ib) { return (a < b) || (a == b && ia < ib); }

// kq[y][x] = k'_b | k'_g<<8 | k'_r<<16 | m'_b<<24,  mg[y][x] = m'_g   (pitch Wp, pad columns zero)
//   k' = k - kmin (bgdehaze/main.py:17 numerator), m' = window min - kmin, 0 where the zero padding of
//   transmission_map (BGDehaze.py:32) reaches into the window.
__global__ void __launch_bounds__(WK_THREADS) window_kernel(const uint8_t* __restrict__ src, int W, int H, int Wp, int wmax,
                                                            const FrameState* __restrict__ fs, uint32_t* __restrict__ kq,
                                                            uint8_t* __restrict__ mgp, ArgPartial* __restrict__ partials) {
  extern __shared__ uint32_t s_w[];
  __shared__ double s_nrm[256];
  __shared__ ArgPartial s_part[WK_THREADS / 32];
  const int wmin = WK_TWIN;
  const int pmax = wmax / 2, pmin = wmin / 2;
  const int PL = max(pmax, pmin), PR = max(wmax - 1 - pmax, wmin - 1 - pmin);
  const int RW = WK_TX + PL + PR, RH = WK_TY + PL + PR;
  uint32_t* P0 = s_w;                 // (B | G<<16) of the input region, replicate border
  uint32_t* P1 = P0 + RW * RH;        // R
  uint32_t* HX0 = P1 + RW * RH;       // horizontal max (B,G)   [RH][WK_TX]
  uint32_t* HX1 = HX0 + RH * WK_TX;   // horizontal max R
  uint32_t* HN0 = HX1 + RH * WK_TX;   // horizontal min (B,G)
  int f = blockIdx.z;
  const uint8_t* img = src + (size_t)f * W * H * 3;
  int x0 = blockIdx.x * WK_TX, y0 = blockIdx.y * WK_TY;
  int kmin = fs[f].kmin, range = (int)fs[f].kmax - kmin;
  s_nrm[threadIdx.x] = (double)threadIdx.x / (double)range;  // normI value of k' (main.py:17)
  for (int i = threadIdx.x; i < RW * RH; i += WK_THREADS) {
    int ry = i / RW, rx = i - ry * RW;
    int y = min(max(y0 - PL + ry, 0), H - 1), x = min(max(x0 - PL + rx, 0), W - 1);
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    P0[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 16);
    P1[i] = p[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RH * WK_TX; i += WK_THREADS) {
    int ry = i / WK_TX, x = i - ry * WK_TX;
    const uint32_t* r0 = P0 + ry * RW + x + PL;
    const uint32_t* r1 = P1 + ry * RW + x + PL;
    uint32_t m0 = 0, m1 = 0, n0 = 0xffffffffu;
    for (int k = -pmax; k < wmax - pmax; k++) {
      m0 = __vmaxu2(m0, r0[k]);
      m1 = max(m1, r1[k]);
    }
    for (int k = -pmin; k < wmin - pmin; k++) n0 = __vminu2(n0, r0[k]);
    HX0[i] = m0; HX1[i] = m1; HN0[i] = n0;
  }
  __syncthreads();
  int x = threadIdx.x % WK_TX, rg = threadIdx.x / WK_TX;
  double bd0 = 0, bd1 = 0;
  unsigned bi0 = 0xffffffffu, bi1 = 0xffffffffu;
  bool have = false;
  int gx = x0 + x;
  uint32_t* kqf = kq + (size_t)f * Wp * H;
  uint8_t* mgf = mgp + (size_t)f * Wp * H;
  for (int j = 0; j < WK_TY / (WK_THREADS / WK_TX); j++) {
    int yy = rg * (WK_TY / (WK_THREADS / WK_TX)) + j;
    int gy = y0 + yy;
    if (gx >= Wp || gy >= H) continue;
    size_t ppix = (size_t)gy * Wp + gx;
    if (gx >= W) { kqf[ppix] = 0u; mgf[ppix] = 0; continue; }
    uint32_t m0 = 0, m1 = 0, n0 = 0xffffffffu;
    for (int k = -pmax; k < wmax - pmax; k++) {
      m0 = __vmaxu2(m0, HX0[(yy + PL + k) * WK_TX + x]);
      m1 = max(m1, HX1[(yy + PL + k) * WK_TX + x]);
    }
    for (int k = -pmin; k < wmin - pmin; k++) n0 = __vminu2(n0, HN0[(yy + PL + k) * WK_TX + x]);
    const int pw = WK_TWIN / 2;
    bool touches = (gx < pw) || (gy < pw) || (gx - pw + WK_TWIN - 1 >= W) || (gy - pw + WK_TWIN - 1 >= H);
    uint32_t c0 = P0[(yy + PL) * RW + x + PL], c1 = P1[(yy + PL) * RW + x + PL];
    uint32_t mb = touches ? 0u : (n0 & 0xffffu) - (uint32_t)kmin;
    uint32_t mg = touches ? 0u : (n0 >> 16) - (uint32_t)kmin;
    kqf[ppix] = ((c0 & 0xffffu) - kmin) | (((c0 >> 16) - kmin) << 8) | ((c1 - kmin) << 16) | (mb << 24);
    mgf[ppix] = (uint8_t)mg;
    // D (BGDehaze.py:20-21): max_R - max_B, max_R - max_G on the normalised image, in fp64
    double nr = s_nrm[(int)m1 - kmin];
    double d0 = nr - s_nrm[(int)(m0 & 0xffffu) - kmin];
    double d1 = nr - s_nrm[(int)(m0 >> 16) - kmin];
    unsigned idx = (unsigned)((size_t)gy * W + gx);
    if (!have) { bd0 = d0; bd1 = d1; bi0 = bi1 = idx; have = true; }
    else {
      if (d0 < bd0) { bd0 = d0; bi0 = idx; }
      if (d1 < bd1) { bd1 = d1; bi1 = idx; }
    }
  }
  // block reduction, lexicographic (value, flat index): first index of the minimum
  if (!have) { bd0 = bd1 = __longlong_as_double(0x7ff0000000000000ll); }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    double o0 = __shfl_xor_sync(0xffffffffu, bd0, d), o1 = __shfl_xor_sync(0xffffffffu, bd1, d);
    unsigned j0 = __shfl_xor_sync(0xffffffffu, bi0, d), j1 = __shfl_xor_sync(0xffffffffu, bi1, d);
    if (lex_less(o0, j0, bd0, bi0)) { bd0 = o0; bi0 = j0; }
    if (lex_less(o1, j1, bd1, bi1)) { bd1 = o1; bi1 = j1; }
  }
  if ((threadIdx.x & 31) == 0) { ArgPartial a; a.d0 = bd0; a.d1 = bd1; a.i0 = bi0; a.i1 = bi1; s_part[threadIdx.x >> 5] = a; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ArgPartial a = s_part[0];
    for (int k = 1; k < WK_THREADS / 32; k++) {
      ArgPartial b = s_part[k];
      if (lex_less(b.d0, b.i0, a.d0, a.i0)) { a.d0 = b.d0; a.i0 = b.i0; }
      if (lex_less(b.d1, b.i1, a.d1, a.i1)) { a.d1 = b.d1; a.i1 = b.i1; }
    }
    partials[(size_t)f * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x] = a;
  }
}

// ---- fast path: both windows 15x15 (the reference's only reachable configuration unless -w is given) ----
// Tile 64 x 32 outputs, 256 threads, three CTAs per SM (64 x 48 at two CTAs per SM measured 8 % slower).  The window max / min of a run of outputs is formed in registers by
// doubling: m2[i] = op(p[i], p[i+1]), m4[i] = op(m2[i], m2[i+2]), m8[i] = op(m4[i], m4[i+4]),
// m15[i] = op(m8[i], m8[i+7])  (4 ops per output instead of 14), on 16-bit pairs (VIMNMX.U16x2).
constexpr int WF_TX = 64, WF_TY = 32, WF_R = 7, WF_RH = WF_TY + 2 * WF_R;   // region rows
constexpr int WF_VR = WF_TY / 4, WF_CTAS = 3;  // output rows per vertical run (four runs of 64 columns); CTAs per SM
constexpr int WF_PP = 84;   // region pitch in words: 78 used; 21 x 16 bytes (odd) -> conflict-free 16-byte loads down a column of rows
constexpr int WF_HP = 68;   // pitch of the horizontal results: 17 x 16 bytes

template <int N, bool MAX>
__device__ __forceinline__ void win15(const uint32_t (&p)[N + 14], uint32_t (&o)[N]) {
  uint32_t a[N + 13], b[N + 11], c[N + 7];
#pragma unroll
  for (int i = 0; i < N + 13; i++) a[i] = MAX ? __vmaxu2(p[i], p[i + 1]) : __vminu2(p[i], p[i + 1]);
#pragma unroll
  for (int i = 0; i < N + 11; i++) b[i] = MAX ? __vmaxu2(a[i], a[i + 2]) : __vminu2(a[i], a[i + 2]);
#pragma unroll
  for (int i = 0; i < N + 7; i++) c[i] = MAX ? __vmaxu2(b[i], b[i + 4]) : __vminu2(b[i], b[i + 4]);
#pragma unroll
  for (int i = 0; i < N; i++) o[i] = MAX ? __vmaxu2(c[i], c[i + 7]) : __vminu2(c[i], c[i + 7]);
}

// candidate for the arg-min of D: exact integer difference first (it decides whenever it differs: one
// unit is 1/range, the fp64 rounding of BGDehaze.py:20-21 is far below that), then the reference's
// fp64 value, then the flat index
struct ArgCand { int di; double d; unsigned idx; };
__device__ __forceinline__ bool cand_less(int ai, double ad, unsigned aidx, int bi, double bd, unsigned bidx) {
  return (ai < bi) || (ai == bi && ((ad < bd) || (ad == bd && aidx < bidx)));
}

__global__ void __launch_bounds__(256, WF_CTAS) window15_kernel(const uint8_t* __restrict__ src, int W, int H, int Wp,
                                                          const FrameState* __restrict__ fs, uint32_t* __restrict__ kq,
                                                          uint8_t* __restrict__ mgp, ArgPartial* __restrict__ partials) {
  extern __shared__ __align__(16) uint32_t s_w[];
  __shared__ double s_nrm[256];
  __shared__ ArgPartial s_part[8];
  uint32_t* P0 = s_w;                       // [WF_RH][84] (B | G<<16), replicate border
  uint32_t* P1 = P0 + WF_RH * WF_PP;        // [WF_RH][84] R
  uint32_t* HX0 = P1 + WF_RH * WF_PP;       // [WF_RH][68] horizontal max (B,G)
  uint32_t* HX1 = HX0 + WF_RH * WF_HP;      // horizontal max R
  uint32_t* HN0 = HX1 + WF_RH * WF_HP;      // horizontal min (B,G)
  const int f = blockIdx.z, tid = threadIdx.x;
  const uint8_t* img = src + (size_t)f * W * H * 3;
  const int x0 = blockIdx.x * WF_TX, y0 = blockIdx.y * WF_TY;
  const int kmin = fs[f].kmin, range = (int)fs[f].kmax - kmin;
  s_nrm[tid] = (double)tid / (double)range;  // normI value of k' (main.py:17)
  // region [y0-7, y0+WF_TY+7) x [x0-7, x0+73) : WF_RH x 80 pixels (78 used + 2 for the vector loads).
  // thread = (column, row mod 3); the loads of several rows are issued before the first use
  const bool words = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 3) == 0);  // rows start on 4-byte boundaries
  if (words) {
    // the 240 bytes of a region row come in as aligned 32-bit words (one coalesced request per row instead of 240 byte
    // loads) into a raw staging area that borrows HX0 (free until the horizontal phase), then get unpacked from shared
    uint32_t* RAW = HX0;  // [WF_RH][64] words
    const int xa = max(x0 - WF_R, 0), xb = min(x0 - WF_R + 79, W - 1);  // first / last image column of the region
    const int byte0 = (xa * 3) & ~3;
    const int nw = ((xb * 3 + 2) >> 2) - (byte0 >> 2) + 1;             // <= 61
#pragma unroll 4
    for (int idx = tid; idx < WF_RH * 64; idx += 256) {
      const int ry = idx >> 6, wi = idx & 63;
      if (wi < nw) {
        const int y = min(max(y0 - WF_R + ry, 0), H - 1);
        RAW[idx] = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)y * W * 3 + byte0) + wi);
      }
    }
    __syncthreads();
    if (tid < 240) {
      const int rx = tid % 80, rr = tid / 80;
      const int xc = min(max(x0 - WF_R + rx, 0), W - 1);
      const uint8_t* rawb = reinterpret_cast<const uint8_t*>(RAW) + (xc * 3 - byte0);
#pragma unroll 7
      for (int ry = rr; ry < WF_RH; ry += 3) {
        const uint8_t* p = rawb + ry * 256;
        uint32_t b = p[0], g = p[1], r = p[2];
        P0[ry * WF_PP + rx] = b | (g << 16);
        P1[ry * WF_PP + rx] = r;
      }
    }
  } else if (tid < 240) {
    const int rx = tid % 80, rr = tid / 80;
    const int xc = min(max(x0 - WF_R + rx, 0), W - 1);
    const uint8_t* col = img + (size_t)xc * 3;
#pragma unroll 7
    for (int ry = rr; ry < WF_RH; ry += 3) {
      const int y = min(max(y0 - WF_R + ry, 0), H - 1);
      const uint8_t* p = col + (size_t)y * W * 3;
      uint32_t b = __ldg(p), g = __ldg(p + 1), r = __ldg(p + 2);
      P0[ry * WF_PP + rx] = b | (g << 16);
      P1[ry * WF_PP + rx] = r;
    }
  }
  __syncthreads();
  // horizontal: task = (row, run of 8 outputs); consecutive lanes take consecutive rows
  for (int task = tid; task < WF_RH * 8; task += 256) {
    int ry = task % WF_RH, j = task / WF_RH;
    const uint4* r0 = reinterpret_cast<const uint4*>(P0 + ry * WF_PP + 8 * j);
    const uint4* r1 = reinterpret_cast<const uint4*>(P1 + ry * WF_PP + 8 * j);
    uint32_t p[22], o[8];
#pragma unroll
    for (int q = 0; q < 6; q++) {
      uint4 v = r0[q];
      if (4 * q < 22) p[4 * q] = v.x;
      if (4 * q + 1 < 22) p[4 * q + 1] = v.y;
      if (4 * q + 2 < 22) p[4 * q + 2] = v.z;
      if (4 * q + 3 < 22) p[4 * q + 3] = v.w;
    }
    win15<8, true>(p, o);
    uint4* h0 = reinterpret_cast<uint4*>(HX0 + ry * WF_HP + 8 * j);
    h0[0] = make_uint4(o[0], o[1], o[2], o[3]); h0[1] = make_uint4(o[4], o[5], o[6], o[7]);
    win15<8, false>(p, o);
    uint4* hn = reinterpret_cast<uint4*>(HN0 + ry * WF_HP + 8 * j);
    hn[0] = make_uint4(o[0], o[1], o[2], o[3]); hn[1] = make_uint4(o[4], o[5], o[6], o[7]);
#pragma unroll
    for (int q = 0; q < 6; q++) {
      uint4 v = r1[q];
      if (4 * q < 22) p[4 * q] = v.x;
      if (4 * q + 1 < 22) p[4 * q + 1] = v.y;
      if (4 * q + 2 < 22) p[4 * q + 2] = v.z;
      if (4 * q + 3 < 22) p[4 * q + 3] = v.w;
    }
    win15<8, true>(p, o);
    uint4* h1 = reinterpret_cast<uint4*>(HX1 + ry * WF_HP + 8 * j);
    h1[0] = make_uint4(o[0], o[1], o[2], o[3]); h1[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
  __syncthreads();
  // vertical: thread = (column, run of WF_VR output rows)
  const int x = tid & 63, run = tid >> 6;
  const int gx = x0 + x;
  ArgCand b0 = {0x7fffffff, 0.0, 0xffffffffu}, b1 = b0;
  int bp0 = -1, bp1 = -1;  // (max_R, max_X) pair of the current best: the same pair further down can never win (same D, larger index)
  if (gx < Wp) {
    uint32_t p[WF_VR + 14], mx0[WF_VR], mx1[WF_VR], mn0[WF_VR];
    const int ry0 = run * WF_VR;
#pragma unroll
    for (int i = 0; i < WF_VR + 14; i++) p[i] = HX0[(ry0 + i) * WF_HP + x];
    win15<WF_VR, true>(p, mx0);
#pragma unroll
    for (int i = 0; i < WF_VR + 14; i++) p[i] = HX1[(ry0 + i) * WF_HP + x];
    win15<WF_VR, true>(p, mx1);
#pragma unroll
    for (int i = 0; i < WF_VR + 14; i++) p[i] = HN0[(ry0 + i) * WF_HP + x];
    win15<WF_VR, false>(p, mn0);
    uint32_t* kqf = kq + (size_t)f * Wp * H;
    uint8_t* mgf = mgp + (size_t)f * Wp * H;
    const bool xt = (gx < WF_R) || (gx + WF_R >= W);
#pragma unroll
    for (int i = 0; i < WF_VR; i++) {
      const int gy = y0 + ry0 + i;
      if (gy >= H) break;
      const size_t ppix = (size_t)gy * Wp + gx;
      if (gx >= W) { kqf[ppix] = 0u; mgf[ppix] = 0; continue; }
      const bool touches = xt || (gy < WF_R) || (gy + WF_R >= H);
      const uint32_t c0 = P0[(ry0 + i + WF_R) * WF_PP + x + WF_R], c1 = P1[(ry0 + i + WF_R) * WF_PP + x + WF_R];
      const uint32_t mb = touches ? 0u : (mn0[i] & 0xffffu) - (uint32_t)kmin;
      const uint32_t mg = touches ? 0u : (mn0[i] >> 16) - (uint32_t)kmin;
      kqf[ppix] = ((c0 & 0xffffu) - kmin) | (((c0 >> 16) - kmin) << 8) | ((c1 - kmin) << 16) | (mb << 24);
      mgf[ppix] = (uint8_t)mg;
      // D (BGDehaze.py:20-21): max_R - max_B, max_R - max_G on the normalised image
      const int mr = (int)mx1[i], mB = (int)(mx0[i] & 0xffffu), mG = (int)(mx0[i] >> 16);
      const unsigned idx = (unsigned)((size_t)gy * W + gx);
      const int d0i = mr - mB, d1i = mr - mG;
      const int pr0 = (mr << 8) | mB, pr1 = (mr << 8) | mG;
      if (d0i <= b0.di && pr0 != bp0) {
        double d = s_nrm[mr - kmin] - s_nrm[mB - kmin];
        if (d0i < b0.di || d < b0.d) { b0.di = d0i; b0.d = d; b0.idx = idx; bp0 = pr0; }  // rows go down: idx only grows
      }
      if (d1i <= b1.di && pr1 != bp1) {
        double d = s_nrm[mr - kmin] - s_nrm[mG - kmin];
        if (d1i < b1.di || d < b1.d) { b1.di = d1i; b1.d = d; b1.idx = idx; bp1 = pr1; }
      }
    }
  }
  // block reduction, lexicographic: first index of the minimum
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    int oi0 = __shfl_xor_sync(0xffffffffu, b0.di, d), oi1 = __shfl_xor_sync(0xffffffffu, b1.di, d);
    double o0 = __shfl_xor_sync(0xffffffffu, b0.d, d), o1 = __shfl_xor_sync(0xffffffffu, b1.d, d);
    unsigned j0 = __shfl_xor_sync(0xffffffffu, b0.idx, d), j1 = __shfl_xor_sync(0xffffffffu, b1.idx, d);
    if (cand_less(oi0, o0, j0, b0.di, b0.d, b0.idx)) { b0.di = oi0; b0.d = o0; b0.idx = j0; }
    if (cand_less(oi1, o1, j1, b1.di, b1.d, b1.idx)) { b1.di = oi1; b1.d = o1; b1.idx = j1; }
  }
  __shared__ int s_di[8][2];
  if ((tid & 31) == 0) {
    ArgPartial a; a.d0 = b0.d; a.d1 = b1.d; a.i0 = b0.idx; a.i1 = b1.idx;
    s_part[tid >> 5] = a; s_di[tid >> 5][0] = b0.di; s_di[tid >> 5][1] = b1.di;
  }
  __syncthreads();
  if (tid == 0) {
    ArgPartial a = s_part[0];
    int i0 = s_di[0][0], i1 = s_di[0][1];
    for (int k = 1; k < 8; k++) {
      ArgPartial b = s_part[k];
      if (cand_less(s_di[k][0], b.d0, b.i0, i0, a.d0, a.i0)) { i0 = s_di[k][0]; a.d0 = b.d0; a.i0 = b.i0; }
      if (cand_less(s_di[k][1], b.d1, b.i1, i1, a.d1, a.i1)) { i1 = s_di[k][1]; a.d1 = b.d1; a.i1 = b.i1; }
    }
    // empty tile part or NaN D (range 0): leave the "nothing found" marker for bglight_finish_kernel
    if (!(a.d0 == a.d0)) a.i0 = 0xffffffffu;
    if (!(a.d1 == a.d1)) a.i1 = 0xffffffffu;
    if (a.i0 == 0xffffffffu) a.d0 = __longlong_as_double(0x7ff0000000000000ll);
    if (a.i1 == 0xffffffffu) a.d1 = __longlong_as_double(0x7ff0000000000000ll);
    partials[(size_t)f * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x] = a;
  }
}

// finish the arg-min and form the background light B = mean of the two selected pixels (BGDehaze.py:22-26)
__global__ void __launch_bounds__(256) bglight_finish_kernel(const uint8_t* __restrict__ src, int W, int H,
                                                             const ArgPartial* __restrict__ partials, int n_part, FrameState* fs,
                                                             int write_B, int write_Bt) {
  __shared__ ArgPartial s_part[8];
  int f = blockIdx.x;
  const ArgPartial* pp = partials + (size_t)f * n_part;
  double bd0 = __longlong_as_double(0x7ff0000000000000ll), bd1 = bd0;
  unsigned bi0 = 0xffffffffu, bi1 = 0xffffffffu;
  for (int i = threadIdx.x; i < n_part; i += 256) {
    ArgPartial a = pp[i];
    if (lex_less(a.d0, a.i0, bd0, bi0)) { bd0 = a.d0; bi0 = a.i0; }
    if (lex_less(a.d1, a.i1, bd1, bi1)) { bd1 = a.d1; bi1 = a.i1; }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    double o0 = __shfl_xor_sync(0xffffffffu, bd0, d), o1 = __shfl_xor_sync(0xffffffffu, bd1, d);
    unsigned j0 = __shfl_xor_sync(0xffffffffu, bi0, d), j1 = __shfl_xor_sync(0xffffffffu, bi1, d);
    if (lex_less(o0, j0, bd0, bi0)) { bd0 = o0; bi0 = j0; }
    if (lex_less(o1, j1, bd1, bi1)) { bd1 = o1; bi1 = j1; }
  }
  if ((threadIdx.x & 31) == 0) { ArgPartial a; a.d0 = bd0; a.d1 = bd1; a.i0 = bi0; a.i1 = bi1; s_part[threadIdx.x >> 5] = a; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ArgPartial a = s_part[0];
    for (int k = 1; k < 8; k++) {
      ArgPartial b = s_part[k];
      if (lex_less(b.d0, b.i0, a.d0, a.i0)) { a.d0 = b.d0; a.i0 = b.i0; }
      if (lex_less(b.d1, b.i1, a.d1, a.i1)) { a.d1 = b.d1; a.i1 = b.i1; }
    }
    FrameState& s = fs[f];
    int kmin = s.kmin;
    double range = (double)((int)s.kmax - kmin);
    if (a.i0 == 0xffffffffu || a.i1 == 0xffffffffu) {  // constant frame: every D is NaN; np.argmin -> 0
      a.i0 = a.i1 = 0;
      s.nan_flag = 1;
    }
    if (write_B) { s.idx0 = a.i0; s.idx1 = a.i1; }
    const uint8_t* img = src + (size_t)f * W * H * 3;
    for (int c = 0; c < 3; c++) {
      double v0 = (double)((int)img[(size_t)a.i0 * 3 + c] - kmin) / range;
      double v1 = (double)((int)img[(size_t)a.i1 * 3 + c] - kmin) / range;
      double b = (v0 + v1) / 2.0;
      if (write_B) s.B[c] = b;
      if (write_Bt) s.Bt[c] = b;
    }
  }
}

// transmission_map output for the stage-wise API: t = 1 - min_window(I_c / B_c), zero padded
__global__ void traw_kernel(const uint32_t* __restrict__ kq, const uint8_t* __restrict__ mgp, int W, int H, int Wp,
                            const FrameState* __restrict__ fs, double* __restrict__ t_raw) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const FrameState& s = fs[0];
  double range = (double)((int)s.kmax - (int)s.kmin);
  size_t pp = (size_t)y * Wp + x, pix = (size_t)y * W + x;
  int mb = (int)(kq[pp] >> 24), mg = (int)mgp[pp];
  t_raw[pix] = 1.0 - ((double)mb / range) / s.Bt[0];
  t_raw[(size_t)W * H + pix] = 1.0 - ((double)mg / range) / s.Bt[1];
}

// ------------------------------------------------------------------------------------------------
// guided-filter helpers
// ------------------------------------------------------------------------------------------------
// 3x3 symmetric solve shared by GF1a / GF2a.  Inputs are window sums in k units:
//   S[3] (sum k_i), SS[6] (sum k_i k_j: 00 01 02 11 12 22), N (window pixel count),
//   Sp (sum p), Sip[3] (sum k_i p).  Output a[3] (k units) and b.
__device__ __forceinline__ void gf_build_M(const uint32_t* si, double N, double epsN2, double* M, double* Sd) {
  Sd[0] = u2d(si[0]); Sd[1] = u2d(si[1]); Sd[2] = u2d(si[2]);
  // exact: N*S_ij and S_i*S_j are integers < 2^53
  M[0] = fma(N, u2d(si[3]), -Sd[0] * Sd[0]) + epsN2;
  M[1] = fma(N, u2d(si[4]), -Sd[0] * Sd[1]);
  M[2] = fma(N, u2d(si[5]), -Sd[0] * Sd[2]);
  M[3] = fma(N, u2d(si[6]), -Sd[1] * Sd[1]) + epsN2;
  M[4] = fma(N, u2d(si[7]), -Sd[1] * Sd[2]);
  M[5] = fma(N, u2d(si[8]), -Sd[2] * Sd[2]) + epsN2;
}
// adjugate of the symmetric matrix [[m0 m1 m2],[m1 m3 m4],[m2 m4 m5]] and 1/det
__device__ __forceinline__ void gf_adjugate(const double* M, double* A, double& rdet) {
  A[0] = M[3] * M[5] - M[4] * M[4];
  A[1] = M[2] * M[4] - M[1] * M[5];
  A[2] = M[1] * M[4] - M[2] * M[3];
  A[3] = M[0] * M[5] - M[2] * M[2];
  A[4] = M[1] * M[2] - M[0] * M[4];
  A[5] = M[0] * M[3] - M[1] * M[1];
  double det = M[0] * A[0] + M[1] * A[1] + M[2] * A[2];
  rdet = rcp_fast(det);
}
__device__ __forceinline__ void gf_solve(const double* A, double rdet, const double* Sd, double N, double invN, double Sp,
                                         const double* Sip, double* a, double& b) {
  // C_i = N*S_ip - S_i*S_p  (= N^2 cov_k);  a_k = C * adj(M) / det(M)
  double C0 = fma(N, Sip[0], -Sd[0] * Sp), C1 = fma(N, Sip[1], -Sd[1] * Sp), C2 = fma(N, Sip[2], -Sd[2] * Sp);
  a[0] = (C0 * A[0] + C1 * A[1] + C2 * A[2]) * rdet;
  a[1] = (C0 * A[1] + C1 * A[3] + C2 * A[4]) * rdet;
  a[2] = (C0 * A[2] + C1 * A[4] + C2 * A[5]) * rdet;
  b = (Sp - a[0] * Sd[0] - a[1] * Sd[1] - a[2] * Sd[2]) * invN;
}

// one coefficient chunk (a0, a1, a2, b) as stored: on the 2^-34 grid (absolute step 6e-11), then f32
__device__ __forceinline__ float4 coef_pack(const double* a, double b) {
  return make_float4((float)grid_round<34>(a[0]), (float)grid_round<34>(a[1]), (float)grid_round<34>(a[2]), (float)grid_round<34>(b));
}

struct GfCommon {
  const uint32_t* kq;     // [n][H][Wp] packed k'_b k'_g k'_r m'_b
  const uint8_t* mg;      // [n][H][Wp] m'_g
  uint32_t* ycc;          // [n][H][Wp] packed Yi Cri Cbi Yj
  const double* stab;     // [n][256][256] exposure ratio S(yi', yj')
  float* splane;          // [n][H][Wp] S per pixel (f32)
  float* ab;              // [n][8][H][Wp]
  float* J;               // [n][2][H][Wp]
  float* refS;            // [n][H][Wp]
  FrameState* fs;
  double eps, tmin;
  double* dbg_tref;       // optional [2][H*W] (frame 0 only)
};

// strip geometry of one launch (see the file header)
struct GfGeom {
  int W, H, Wp;   // image size; pitch of the internal planes (multiple of 4)
  int r;          // box radius
  int HL;         // halo rounded up to a multiple of 4
  int SW;         // output columns per strip (multiple of 4)
  int NQ;         // quads per strip: one zero guard quad + (2*HL + SW)/4
  int seg_h;      // output rows per vertical segment
  int fast;       // r % 4 == 0: window edges fall on quad boundaries
};

// ---- shared per-CTA frame constants --------------------------------------------------------------
struct FrameConst {
  int kmin, range;
  double B[3], Bt[3];
  // restored-image parameters (valid after GF1b): J min / (max-min)
  double jmin[2], jinv[2], jrcp[2];
  int yi_min, yi_rng, yj_min, yj_rng;
};

__device__ __forceinline__ void load_frame_const(const FrameState& s, FrameConst& c) {
  c.kmin = s.kmin;
  c.range = (int)s.kmax - (int)s.kmin;
  c.B[0] = s.B[0]; c.B[1] = s.B[1]; c.B[2] = s.B[2];
  c.Bt[0] = s.Bt[0]; c.Bt[1] = s.Bt[1]; c.Bt[2] = s.Bt[2];
  for (int k = 0; k < 2; k++) {
    double mn = dunkey(s.jmin_key[k]), mx = dunkey(s.jmax_key[k]);
    c.jmin[k] = mn;
    c.jinv[k] = mx - mn;  // denominator; divisions are done where used
    c.jrcp[k] = 1.0 / (mx - mn);
  }
  c.yi_min = s.yi_min; c.yi_rng = (int)s.yi_max - (int)s.yi_min;
  c.yj_min = s.yj_min; c.yj_rng = (int)s.yj_max - (int)s.yj_min;
}

// red channel of `restored` (RC_correction, BGDehaze.py:61-64) as a function of k'_r, plus the
// truncated bytes R8 / I8 of adaptiveExp_map (BGDehaze.py:75-76).
struct RedTables {
  double redN[256];     // normRrec for k' = 0..255
  uint8_t red8[256];    // (normRrec*255).astype(uint8)
  uint8_t i8[256];      // (normI*255).astype(uint8) for k'
};
__device__ __forceinline__ int trunc_u8(double v) {  // numpy float64 -> uint8 cast for v in [0,255]; NaN -> 0
  if (!(v == v)) return 0;
  int i = (int)v;
  return i & 0xff;
}
__device__ void build_red_tables(const FrameState& s, const FrameConst& fc, double n_px, RedTables* rt, int tid, int nthreads) {
  double range = (double)fc.range;
  double mean_b = ((double)s.jsum_fix[0] * (1.0 / 4294967296.0) / n_px - fc.jmin[0]) / fc.jinv[0];
  double mean_g = ((double)s.jsum_fix[1] * (1.0 / 4294967296.0) / n_px - fc.jmin[1]) / fc.jinv[1];
  double avgRr = 1.5 - mean_b - mean_g;
  double mean_r = ((double)s.rsum / n_px) / range;
  double coef = avgRr / mean_r;
  double ra = ((double)s.rmin / range) * coef, rb = ((double)s.rmax / range) * coef;
  double rmn = fmin(ra, rb), rmx = fmax(ra, rb);
  for (int k = tid; k < 256; k += nthreads) {
    double nk = (double)k / range;
    double v = (nk * coef - rmn) / (rmx - rmn);
    rt->redN[k] = v;
    rt->red8[k] = (uint8_t)trunc_u8(v * 255.0);
    rt->i8[k] = (uint8_t)trunc_u8(nk * 255.0);
  }
}

struct ExpShared {
  RedTables rt;
  FrameConst fc;
};
__device__ void exp_shared_init(ExpShared* sh, const FrameState& s, double n_px) {
  if (threadIdx.x == 0) load_frame_const(s, sh->fc);
  __syncthreads();
  build_red_tables(s, sh->fc, n_px, &sh->rt, threadIdx.x, blockDim.x);
  __syncthreads();
}
// restored blue / green from the stored J value (dehazed_BG, BGDehaze.py:53-56).  The IEEE division is
// what the bytes R8 are truncated from; the reciprocal flavour feeds products only.
__device__ __forceinline__ double norm_j(float j, const FrameConst& fc, int c) { return ((double)j - fc.jmin[c]) / fc.jinv[c]; }
__device__ __forceinline__ double norm_j_fast(float j, const FrameConst& fc, int c) { return ((double)j - fc.jmin[c]) * fc.jrcp[c]; }
// (restored*255).astype(uint8) of BGDehaze.py:75: the byte is a truncation, so the quotient must be the
// IEEE one only where restored*255 sits next to an integer; everywhere else the reciprocal product (a few
// ulp off) truncates to the same byte.
__device__ __forceinline__ int restored_byte(float j, const FrameConst& fc, int c) {
  double v = norm_j_fast(j, fc, c) * 255.0;
  if (fabs(v - rint(v)) < 1.0e-9 || !(v == v)) v = norm_j(j, fc, c) * 255.0;
  return trunc_u8(v);
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d); v = o < v ? o : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
  return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_min_f32(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

__device__ __forceinline__ uint32_t quad_get(const uint4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }
__device__ __forceinline__ float quad_get(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }


// Coefficient rows (the a,b of a guided filter) are stored quad-interleaved: per image row, per quad of
// four pixels, NP/4 sixteen-byte chunks per pixel (chunk = a0,a1,a2,b of one filter) - 128 bytes per quad
// for GF1 (two filters), 64 for GF2.  The chunk order inside a quad is XOR-swizzled with the quad index
// so that the V-phase of the reader (one thread per quad, a 16-byte shared load per chunk) is free of
// bank conflicts although consecutive lanes are a whole quad apart.  Writer and reader are both ours.
template <int NP>
__device__ __forceinline__ int coef_chunk(int gq, int c, int j) {  // physical chunk of (pixel c, filter j) in quad gq
  if (NP == 8) return (2 * c + j) ^ (gq & 7);
  return c ^ ((gq >> 1) & 3);
}

// V-phase of the coefficient readers (GF1b, GF2b): the entering and the leaving row were brought to shared
// memory by one TMA bulk copy each ([2][NT quads][NP chunks]); every value goes through fp64.
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <int NP, int NT, bool ENTER, bool LEAVE, bool FULL>
__device__ __forceinline__ void gf_accum_staged_case(const float4* __restrict__ stg, int gq, unsigned cmask, double (&Vd)[4][NP]) {
  // quad blocks are NP*16 bytes and block-aligned, so (logical chunk ^ swizzle) * 16 is the block address
  // with the swizzle folded in, XOR a compile-time constant: one LOP3 per load
  const unsigned swz = (unsigned)coef_chunk<NP>(gq, 0, 0) << 4;
  const unsigned be = (unsigned)__cvta_generic_to_shared(stg + (size_t)threadIdx.x * NP) + swz;
  const unsigned bl = (unsigned)__cvta_generic_to_shared(stg + (size_t)(NT + threadIdx.x) * NP) + swz;
#pragma unroll
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int j = 0; j < NP / 4; j++) {
      const unsigned q16 = (unsigned)(NP == 8 ? 2 * c + j : c) << 4;
      float4 e, l;
      if (ENTER) e = lds128(be ^ q16);
      if (LEAVE) l = lds128(bl ^ q16);
      if (FULL || (cmask & (1u << c))) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
          if (ENTER && LEAVE) Vd[c][4 * j + m] += (double)quad_get(e, m) - (double)quad_get(l, m);
          else if (ENTER) Vd[c][4 * j + m] += (double)quad_get(e, m);
          else Vd[c][4 * j + m] -= (double)quad_get(l, m);
        }
      }
    }
  }
}
// uniform dispatch: steady state (both rows, all four columns inside the image) is the straight-line case
template <int NP, int NT>
__device__ __forceinline__ void gf_accum_staged(const float4* __restrict__ stg, int gq, bool enter, bool leave, unsigned cmask, double (&Vd)[4][NP]) {
  if (enter && leave) {
    if (cmask == 0xfu) gf_accum_staged_case<NP, NT, true, true, true>(stg, gq, cmask, Vd);
    else gf_accum_staged_case<NP, NT, true, true, false>(stg, gq, cmask, Vd);
  } else if (enter) {
    gf_accum_staged_case<NP, NT, true, false, false>(stg, gq, cmask, Vd);
  } else if (leave) {
    gf_accum_staged_case<NP, NT, false, true, false>(stg, gq, cmask, Vd);
  }
}

// -------------------------------------------------------------------------------------------------
// policies: what is accumulated per pixel (accum), and what is made of the window sums (column)
// -------------------------------------------------------------------------------------------------
constexpr int GF_NSEG = 8;    // scan segments per quad-total row
constexpr int GF_SEGQ = 28;   // quads per segment (7 x 16 bytes: an odd chunk count keeps the vector loads conflict-free)
constexpr int GF_GP = GF_NSEG * GF_SEGQ;  // pitch of a quad-total row = max threads per CTA

// GF1a: guide = normI (k units), p = max(t_blue, tmin) and max(t_green, tmin)
struct PolGF1a {
  static constexpr int NI = 9, ND = 8, MINB = 1, MAXREG = 255, NT = 224, NAUX = 1;
  static constexpr bool PREFETCH = true, INT_HALF = false, DBUF = true, META = true, DELAY = false;
  static constexpr int EARLY_ROW = 0, ROWST_BYTES = 0;
  static constexpr bool EARLY_SCAN = false, UNCOND_STAGE = true;
  struct Shared {
    double pT[2][256];  // p_c as a function of the window-min k'
    FrameConst fc;
    double epsN_k;      // eps * range^2 (read at its use: a register held across the march ends up in a spill slot)
  };
  struct Raw { uint4 k; uint32_t m; };
  GfCommon g; Shared* sh; int Wp, H, f;
  const uint32_t* kq; const uint8_t* mg; float* ab;
  __device__ void init(const GfCommon& gc, int frame, Shared* s, const GfGeom& gg) {
    g = gc; sh = s; Wp = gg.Wp; H = gg.H; f = frame;
    size_t n_pp = (size_t)Wp * H;
    kq = g.kq + (size_t)f * n_pp;
    mg = g.mg + (size_t)f * n_pp;
    ab = g.ab + (size_t)f * 8 * n_pp;
    if (threadIdx.x == 0) load_frame_const(g.fs[f], sh->fc);
    __syncthreads();
    double range = (double)sh->fc.range;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      int c = i >> 8, k = i & 255;
      double t = 1.0 - ((double)k / range) / sh->fc.Bt[c];     // transmission_map (BGDehaze.py:35-36)
      sh->pT[c][k] = grid_round<28>((t < g.tmin) ? g.tmin : t);  // np.maximum(t, tmin) (NaN stays NaN), on the 2^-28 grid
    }
    if (threadIdx.x == 0) sh->epsN_k = g.eps * range * range;
    __syncthreads();
  }
  // staging slot s (0..3 = 2 buffers x enter/leave) of this thread: one uint4 + one u32
  static constexpr int STAGE_BYTES = 4 * NT * 20;
  __device__ __forceinline__ void stage_issue(unsigned char* st, int s, int y, int gx) const {
    size_t o = (size_t)y * Wp + gx;
    cp_async16(st + ((size_t)s * NT + threadIdx.x) * 16, kq + o);
    cp_async4(st + (size_t)4 * NT * 16 + ((size_t)s * NT + threadIdx.x) * 4, mg + o);
  }
  __device__ __forceinline__ void stage_read(const unsigned char* st, int s, Raw& r) const {
    r.k = *reinterpret_cast<const uint4*>(st + ((size_t)s * NT + threadIdx.x) * 16);
    r.m = *reinterpret_cast<const uint32_t*>(st + (size_t)4 * NT * 16 + ((size_t)s * NT + threadIdx.x) * 4);
  }
  template <int SIGN, bool FULL>
  __device__ __forceinline__ void accum(const Raw& r, unsigned cmask, uint32_t (&Vi)[4][NI > 0 ? NI : 1], double (&Vd)[4][ND]) const {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t w = quad_get(r.k, c);
      uint32_t kb = w & 255u, kg = (w >> 8) & 255u, kr = (w >> 16) & 255u, mb = w >> 24, mgv = (r.m >> (8 * c)) & 255u;
      if (SIGN > 0) {
        Vi[c][0] += kb; Vi[c][1] += kg; Vi[c][2] += kr;
        Vi[c][3] += kb * kb; Vi[c][4] += kb * kg; Vi[c][5] += kb * kr;
        Vi[c][6] += kg * kg; Vi[c][7] += kg * kr; Vi[c][8] += kr * kr;
      } else {
        Vi[c][0] -= kb; Vi[c][1] -= kg; Vi[c][2] -= kr;
        Vi[c][3] -= kb * kb; Vi[c][4] -= kb * kg; Vi[c][5] -= kb * kr;
        Vi[c][6] -= kg * kg; Vi[c][7] -= kg * kr; Vi[c][8] -= kr * kr;
      }
      if (FULL || (cmask & (1u << c))) {
        double pb = sh->pT[0][mb], pg = sh->pT[1][mgv];
        if (SIGN < 0) { pb = -pb; pg = -pg; }
        double db = u2d(kb), dg = u2d(kg), dr = u2d(kr);
        Vd[c][0] += pb;
        Vd[c][1] += pg;
        Vd[c][2] = fma(db, pb, Vd[c][2]); Vd[c][3] = fma(dg, pb, Vd[c][3]); Vd[c][4] = fma(dr, pb, Vd[c][4]);
        Vd[c][5] = fma(db, pg, Vd[c][5]); Vd[c][6] = fma(dg, pg, Vd[c][6]); Vd[c][7] = fma(dr, pg, Vd[c][7]);
      }
    }
  }
  __device__ __forceinline__ void row_begin(int, int) {}
  // results go out pixel by pixel: two 16-byte chunks (one per filter) into the swizzled quad block
  __device__ __forceinline__ void column(int, int y, int x, int Ncnt, const uint32_t* si, const double* sd) {
    double N = u2d((uint32_t)Ncnt), invN = rcp_fast(N);
    double M[6], Sd[3], A[6], rdet;
    gf_build_M(si, N, *(volatile double*)&sh->epsN_k * N * N, M, Sd);
    gf_adjugate(M, A, rdet);
    double a[3], b;
    const int gq = x >> 2, c = x & 3;
    float4* blk = reinterpret_cast<float4*>(ab) + ((size_t)y * (Wp >> 2) + gq) * 8;
    gf_solve(A, rdet, Sd, N, invN, sd[0], sd + 2, a, b);
    blk[coef_chunk<8>(gq, c, 0)] = coef_pack(a, b);
    gf_solve(A, rdet, Sd, N, invN, sd[1], sd + 5, a, b);
    blk[coef_chunk<8>(gq, c, 1)] = coef_pack(a, b);
  }
  __device__ __forceinline__ void store_pair(int, int) {}
  __device__ void finish() {}
};

// GF1b: q = (box(a).k + box(b))/N for blue and green -> J (dehazed_BG) + reductions
struct PolGF1b {
  static constexpr int NI = 0, ND = 8, MINB = 1, MAXREG = 255, NT = 224, NAUX = 1;
  static constexpr bool PREFETCH = false, INT_HALF = false, DBUF = true, META = false, DELAY = false;
  static constexpr bool EARLY_SCAN = true, UNCOND_STAGE = false;
  static constexpr int EARLY_ROW = 1, ROWST_BYTES = 0;  // row_begin's one 16-byte load goes out before the accumulate phase
  struct Shared {
    double nrm[256];
    FrameConst fc;
  };
  struct Raw {};
  GfCommon g; Shared* sh; int W, Wp, H, f;
  const uint32_t* kq; const float* ab; float* J;
  float jmn[2], jmx[2];
  long long jsum[2];   // sum of bits(J*2^32 + 1.5*2^52): exact 2^-32 fixed point, independent of the partition
  unsigned cnt, rmn, rmx, rsum, nanf;
  uint4 krow;
  float o[2][2];
  double* dbg;   // stage-wise API only: refined t of frame 0
  __device__ void init(const GfCommon& gc, int frame, Shared* s, const GfGeom& gg) {
    g = gc; sh = s; W = gg.W; Wp = gg.Wp; H = gg.H; f = frame;
    dbg = (frame == 0) ? gc.dbg_tref : nullptr;
    size_t n_pp = (size_t)Wp * H;
    kq = g.kq + (size_t)f * n_pp;
    ab = g.ab + (size_t)f * 8 * n_pp;
    J = g.J + (size_t)f * 2 * n_pp;
    if (threadIdx.x == 0) load_frame_const(g.fs[f], sh->fc);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh->nrm[i] = (double)i / (double)sh->fc.range;
    jmn[0] = jmn[1] = __int_as_float(0x7f800000); jmx[0] = jmx[1] = -__int_as_float(0x7f800000);
    jsum[0] = jsum[1] = 0;
    cnt = 0; rmn = 255; rmx = 0; rsum = 0; nanf = 0;
    __syncthreads();
  }
  static constexpr int NP = 8, STAGE_BYTES = 2 * NP * NT * 16;
  __device__ __forceinline__ const float* coef_rows() const { return ab; }
  __device__ __forceinline__ void row_begin(int y, int gx) { krow = __ldg(reinterpret_cast<const uint4*>(kq + (size_t)y * Wp + gx)); }
  __device__ __forceinline__ void column(int cc, int y, int x, int Ncnt, const uint32_t*, const double* sd) {
    double invN = rcp_fast(u2d((uint32_t)Ncnt));
    uint32_t w = quad_get(krow, x & 3);
    uint32_t k[3] = {w & 255u, (w >> 8) & 255u, (w >> 16) & 255u};
    double kd[3] = {u2d(k[0]), u2d(k[1]), u2d(k[2])};
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double* s = sd + 4 * c;
      double q = (s[0] * kd[0] + s[1] * kd[1] + s[2] * kd[2] + s[3]) * invN;   // guidedfilter.py:100-101
      if (dbg) dbg[(size_t)c * W * H + (size_t)y * W + x] = q;
      double Bc = sh->fc.B[c];
      double Jv = (sh->nrm[k[c]] - Bc) * rcp_fast(q) + Bc;                    // BGDehaze.py:53,55
      float Jf = (float)Jv;
      o[cc][c] = Jf;
      if (!(fabsf(Jf) < 262144.0f)) { nanf |= 1u; Jf = 0.f; }
      jmn[c] = fminf(jmn[c], Jf);
      jmx[c] = fmaxf(jmx[c], Jf);
      jsum[c] += __double_as_longlong(fma((double)Jf, 4294967296.0, 6755399441055744.0));
    }
    cnt++;
    rmn = min(rmn, k[2]); rmx = max(rmx, k[2]); rsum += k[2];
  }
  __device__ __forceinline__ void store_pair(int y, int x) {
    size_t n_pp = (size_t)Wp * H, pp = (size_t)y * Wp + x;
    *reinterpret_cast<float2*>(J + pp) = make_float2(o[0][0], o[1][0]);
    *reinterpret_cast<float2*>(J + n_pp + pp) = make_float2(o[0][1], o[1][1]);
  }
  __device__ void finish() {
    FrameState& s = g.fs[f];
    for (int c = 0; c < 2; c++) {
      float a = warp_min_f32(jmn[c]), b = warp_max_f32(jmx[c]);
      long long sm = warp_sum_i64(jsum[c] - (long long)cnt * __double_as_longlong(6755399441055744.0));
      if ((threadIdx.x & 31) == 0) {
        if (a <= b) {
          atomicMin(&s.jmin_key[c], dkey((double)a));
          atomicMax(&s.jmax_key[c], dkey((double)b));
        }
        atomicAdd((unsigned long long*)&s.jsum_fix[c], (unsigned long long)sm);
      }
    }
    unsigned a = warp_reduce_min_u32(rmn), b = warp_reduce_max_u32(rmx);
    unsigned rs = __reduce_add_sync(0xffffffffu, rsum);
    unsigned nf = __reduce_or_sync(0xffffffffu, nanf);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s.rmin, a);
      atomicMax(&s.rmax, b);
      atomicAdd(&s.rsum, (unsigned long long)rs);
      if (nf) atomicOr(&s.nan_flag, nf);
    }
  }
};

// GF2a: guide = normYiCrCb (k units), p = S (BGDehaze.py:83)
struct PolGF2a {
  static constexpr int NI = 9, ND = 4, MINB = 1, MAXREG = 255, NT = 224, NAUX = 1;
  static constexpr bool PREFETCH = true, INT_HALF = false, DBUF = true, META = true, DELAY = true;
  static constexpr int EARLY_ROW = 0, ROWST_BYTES = 0;
  static constexpr bool EARLY_SCAN = false, UNCOND_STAGE = false;
  struct Shared { FrameConst fc; };
  struct Raw { uint4 y; float4 s; };
  GfCommon g; Shared* sh; int Wp, H, f;
  const uint32_t* ycc; const float* sp; float* ab;
  uint32_t ysub;   // (yi_min, yi_min, yi_min, yj_min): no byte can borrow
  double epsN_k; unsigned nanf;
  __device__ void init(const GfCommon& gc, int frame, Shared* s, const GfGeom& gg) {
    g = gc; sh = s; Wp = gg.Wp; H = gg.H; f = frame;
    size_t n_pp = (size_t)Wp * H;
    ycc = g.ycc + (size_t)f * n_pp;
    sp = g.splane + (size_t)f * n_pp;
    ab = g.ab + (size_t)f * 8 * n_pp;
    if (threadIdx.x == 0) load_frame_const(g.fs[f], sh->fc);
    __syncthreads();
    double rng = (double)sh->fc.yi_rng;
    epsN_k = g.eps * rng * rng;
    uint32_t a = (uint32_t)sh->fc.yi_min, b = (uint32_t)sh->fc.yj_min;
    ysub = a | (a << 8) | (a << 16) | (b << 24);
    nanf = 0;
  }
  // staging slot s of this thread: the packed guide quad and the S quad
  static constexpr int STAGE_BYTES = 4 * NT * 32;
  __device__ __forceinline__ void stage_issue(unsigned char* st, int s, int y, int gx) const {
    size_t o = (size_t)y * Wp + gx;
    cp_async16(st + ((size_t)s * NT + threadIdx.x) * 32, ycc + o);
    cp_async16(st + ((size_t)s * NT + threadIdx.x) * 32 + 16, sp + o);
  }
  __device__ __forceinline__ void stage_read(const unsigned char* st, int s, Raw& r) const {
    r.y = *reinterpret_cast<const uint4*>(st + ((size_t)s * NT + threadIdx.x) * 32);
    r.s = *reinterpret_cast<const float4*>(st + ((size_t)s * NT + threadIdx.x) * 32 + 16);
  }
  template <int SIGN, bool FULL>
  __device__ __forceinline__ void accum(const Raw& r, unsigned cmask, uint32_t (&Vi)[4][NI], double (&Vd)[4][ND]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      if (FULL || (cmask & (1u << c))) {
        uint32_t w = quad_get(r.y, c) - ysub;
        uint32_t g0 = w & 255u, g1 = (w >> 8) & 255u, g2 = (w >> 16) & 255u;
        double S = (double)quad_get(r.s, c);
        if (SIGN > 0) {
          if (!(S == S)) nanf = 1u;
          Vi[c][0] += g0; Vi[c][1] += g1; Vi[c][2] += g2;
          Vi[c][3] += g0 * g0; Vi[c][4] += g0 * g1; Vi[c][5] += g0 * g2;
          Vi[c][6] += g1 * g1; Vi[c][7] += g1 * g2; Vi[c][8] += g2 * g2;
        } else {
          S = -S;
          Vi[c][0] -= g0; Vi[c][1] -= g1; Vi[c][2] -= g2;
          Vi[c][3] -= g0 * g0; Vi[c][4] -= g0 * g1; Vi[c][5] -= g0 * g2;
          Vi[c][6] -= g1 * g1; Vi[c][7] -= g1 * g2; Vi[c][8] -= g2 * g2;
        }
        Vd[c][0] += S;
        Vd[c][1] = fma(u2d(g0), S, Vd[c][1]); Vd[c][2] = fma(u2d(g1), S, Vd[c][2]); Vd[c][3] = fma(u2d(g2), S, Vd[c][3]);
      }
    }
  }
  __device__ __forceinline__ void row_begin(int, int) {}
  __device__ __forceinline__ void column(int, int y, int x, int Ncnt, const uint32_t* si, const double* sd) {
    double N = u2d((uint32_t)Ncnt), invN = rcp_fast(N);
    double M[6], Sd[3], A[6], rdet, a[3], b;
    gf_build_M(si, N, epsN_k * N * N, M, Sd);
    gf_adjugate(M, A, rdet);
    gf_solve(A, rdet, Sd, N, invN, sd[0], sd + 1, a, b);
    const int gq = x >> 2;
    float4* blk = reinterpret_cast<float4*>(ab) + ((size_t)y * (Wp >> 2) + gq) * 4;
    blk[coef_chunk<4>(gq, x & 3, 0)] = coef_pack(a, b);
  }
  __device__ __forceinline__ void store_pair(int, int) {}
  __device__ void finish() {
    unsigned nf = __reduce_or_sync(0xffffffffu, nanf);
    if ((threadIdx.x & 31) == 0 && nf) atomicOr(&g.fs[f].nan_flag, 1u);
  }
};

// GF2b: refined S -> exposure product -> min/max (BGDehaze.py:84-89)
struct PolGF2b {
  static constexpr int NI = 0, ND = 4, MINB = 2, MAXREG = 128, NT = 224, NAUX = 1;
  static constexpr bool PREFETCH = false, INT_HALF = false, DBUF = true, META = false, DELAY = false;
  static constexpr bool EARLY_SCAN = true, UNCOND_STAGE = false;
  static constexpr int EARLY_ROW = 2, ROWST_BYTES = NT * 64;  // row_begin's four 16-byte loads are staged through shared memory (cp.async)
  typedef ExpShared Shared;
  struct Raw {};
  GfCommon g; Shared* sh; int Wp, H, f;
  const uint32_t* kq; const uint32_t* ycc; const float* J; const float* ab; float* refS;
  double omn, omx; unsigned nanf;
  uint4 krow, yrow; float4 jb, jg;
  float o[2];
  __device__ void init(const GfCommon& gc, int frame, Shared* s, const GfGeom& gg) {
    g = gc; sh = s; Wp = gg.Wp; H = gg.H; f = frame;
    size_t n_pp = (size_t)Wp * H;
    kq = g.kq + (size_t)f * n_pp;
    ycc = g.ycc + (size_t)f * n_pp;
    J = g.J + (size_t)f * 2 * n_pp;
    ab = g.ab + (size_t)f * 8 * n_pp;
    refS = g.refS + (size_t)f * n_pp;
    exp_shared_init(sh, g.fs[f], (double)gg.W * (double)gg.H);
    omn = __longlong_as_double(0x7ff0000000000000ll); omx = -omn; nanf = 0;
  }
  static constexpr int NP = 4, STAGE_BYTES = 2 * NP * NT * 16;
  __device__ __forceinline__ const float* coef_rows() const { return ab; }
  __device__ __forceinline__ void row_begin(int y, int gx) {
    size_t n_pp = (size_t)Wp * H, o = (size_t)y * Wp + gx;
    krow = __ldg(reinterpret_cast<const uint4*>(kq + o));
    yrow = __ldg(reinterpret_cast<const uint4*>(ycc + o));
    jb = __ldg(reinterpret_cast<const float4*>(J + o));
    jg = __ldg(reinterpret_cast<const float4*>(J + n_pp + o));
  }
  // the same four loads as asynchronous copies into this thread's staging slots, and their pick-up
  __device__ __forceinline__ void row_prefetch(unsigned char* rs, int y, int gx) const {
    size_t n_pp = (size_t)Wp * H, o = (size_t)y * Wp + gx;
    uint4* slot = reinterpret_cast<uint4*>(rs) + threadIdx.x;
    cp_async16(slot, kq + o);
    cp_async16(slot + NT, ycc + o);
    cp_async16(slot + 2 * NT, J + o);
    cp_async16(slot + 3 * NT, J + n_pp + o);
    cp_async_commit();
  }
  __device__ __forceinline__ void row_pickup(const unsigned char* rs) {
    cp_async_wait_all();
    const uint4* slot = reinterpret_cast<const uint4*>(rs) + threadIdx.x;
    krow = slot[0];
    yrow = slot[NT];
    uint4 a = slot[2 * NT], b = slot[3 * NT];
    jb = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
    jg = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
  }
  __device__ __forceinline__ void column(int cc, int, int x, int Ncnt, const uint32_t*, const double* sd) {
    double invN = rcp_fast(u2d((uint32_t)Ncnt));
    int c4 = x & 3;
    uint32_t yw = quad_get(yrow, c4);
    uint32_t ymin = (uint32_t)sh->fc.yi_min;
    double g0 = u2d((yw & 255u) - ymin), g1 = u2d(((yw >> 8) & 255u) - ymin), g2 = u2d(((yw >> 16) & 255u) - ymin);
    double q = (sd[0] * g0 + sd[1] * g1 + sd[2] * g2 + sd[3]) * invN;
    float qf = (float)q;
    o[cc] = qf;
    double qr = (double)qf;
    double rb = norm_j_fast(quad_get(jb, c4), sh->fc, 0), rg = norm_j_fast(quad_get(jg, c4), sh->fc, 1);
    double rr = sh->rt.redN[(quad_get(krow, c4) >> 16) & 255u];
    // min / max over the three channels of restored * refinedS
    double o0 = rb * qr, o1 = rg * qr, o2 = rr * qr;
    double os = (o0 + o1) + o2;
    if (!(os == os) || fabs(os) > 1.0e300) { nanf = 1u; }  // any NaN / inf poisons the sum
    else {  // plain compare-selects: no NaN in here, fmin/fmax would pay for their NaN rules
      double lo = o0 < o1 ? o0 : o1, hi = o0 < o1 ? o1 : o0;
      lo = o2 < lo ? o2 : lo; hi = o2 > hi ? o2 : hi;
      omn = lo < omn ? lo : omn;
      omx = hi > omx ? hi : omx;
    }
  }
  __device__ __forceinline__ void store_pair(int y, int x) {
    *reinterpret_cast<float2*>(refS + (size_t)y * Wp + x) = make_float2(o[0], o[1]);
  }
  __device__ void finish() {
    double a = warp_min_f64(omn), b = warp_max_f64(omx);
    unsigned nf = __reduce_or_sync(0xffffffffu, nanf);
    if ((threadIdx.x & 31) == 0) {
      if (a <= b) {
        atomicMin(&g.fs[f].omin_key, dkey(a));
        atomicMax(&g.fs[f].omax_key, dkey(b));
      }
      if (nf) atomicOr(&g.fs[f].nan_flag, 1u);
    }
  }
};

// -------------------------------------------------------------------------------------------------
// the quad-march kernel
// -------------------------------------------------------------------------------------------------
// One scan task: GF_SEGQ consecutive quad totals of one moment -> inclusive prefix over the whole row.
// The eight tasks of a moment sit in eight adjacent lanes; their segment totals are exchanged with
// shuffles.  The segment lives in registers between the load and the store (one pass over shared
// memory); the in-register prefix is done per group of four to keep the dependent chain short.
template <class T, class V4>
__device__ __forceinline__ void gf_scan_task(T* row, int seg, bool live) {
  constexpr int VW = sizeof(V4) / sizeof(T);  // 4 (u32) or 2 (f64)
  constexpr int NG = GF_SEGQ / 4;             // groups of four values
  T* p = row + seg * GF_SEGQ;
  T v[GF_SEGQ];
#pragma unroll
  for (int i = 0; i < GF_SEGQ; i += VW) {
    V4 q;
    if (live) q = *reinterpret_cast<const V4*>(p + i);
    if constexpr (VW == 4) { v[i] = live ? q.x : T(0); v[i + 1] = live ? q.y : T(0); v[i + 2] = live ? q.z : T(0); v[i + 3] = live ? q.w : T(0); }
    else { v[i] = live ? q.x : T(0); v[i + 1] = live ? q.y : T(0); }
  }
  T gt[NG];
#pragma unroll
  for (int g = 0; g < NG; g++) {  // prefix inside each group of four (independent chains of three)
    v[4 * g + 1] += v[4 * g]; v[4 * g + 2] += v[4 * g + 1]; v[4 * g + 3] += v[4 * g + 2];
    gt[g] = v[4 * g + 3];
  }
#pragma unroll
  for (int g = 1; g < NG; g++) gt[g] += gt[g - 1];  // inclusive prefix of the group totals
  T incl = gt[NG - 1];
#pragma unroll
  for (int d = 1; d < GF_NSEG; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, incl, d, GF_NSEG);
    if (seg >= d) incl += o;
  }
  T run = __shfl_up_sync(0xffffffffu, incl, 1, GF_NSEG);  // exclusive offset of this segment
  if (seg == 0) run = 0;
  if (live) {
#pragma unroll
    for (int i = 0; i < GF_SEGQ; i += VW) {
      T off = run + (i >= 4 ? gt[i / 4 - 1] : T(0));
      V4 q;
      if constexpr (VW == 4) { q.x = v[i] + off; q.y = v[i + 1] + off; q.z = v[i + 2] + off; q.w = v[i + 3] + off; }
      else { q.x = v[i] + off; q.y = v[i + 1] + off; }
      *reinterpret_cast<V4*>(p + i) = q;
    }
  }
}

// shared-memory layout of one published row; every pitch is a compile-time constant so that each
// access is base register + immediate:
//   Pd01 [ND][NT] double2 (prefix 0,1 of the quad) | Pd2 [ND][NT] f64 (prefix 2; prefix 3 = the quad total
//   is the difference of two neighbouring entries of the scanned totals and is not stored)
//   Gd [ND][GF_GP] f64 | Pi [NI][NT] uint4 | Gi [NI][GF_GP] u32 | policy tables | row staging | mbarrier
template <class P>
struct GfSmem {
  static constexpr int NT = P::NT, NI = P::NI, ND = P::ND;
  static constexpr size_t off_d23 = (size_t)ND * NT * 16;
  static constexpr size_t off_gd = off_d23 + (size_t)ND * NT * 8;
  static constexpr size_t off_pi = off_gd + (size_t)ND * GF_GP * 8;
  static constexpr size_t off_gi = off_pi + (size_t)NI * NT * 16;
  static constexpr size_t buf_bytes = (off_gi + (size_t)NI * GF_GP * 4 + 127) & ~(size_t)127;  // one published row
  // DBUF: two copies addressed by the parity of the output row - the next row can be published while slow
  // warps still read this one, which removes the third CTA barrier of a row
  static constexpr size_t off_sh = buf_bytes * (P::DBUF ? 2 : 1);
  static constexpr size_t off_st = (off_sh + sizeof(typename P::Shared) + 127) & ~(size_t)127;  // quad blocks stay block-aligned
  static constexpr size_t off_bar = off_st + P::STAGE_BYTES;
  static constexpr size_t off_row = off_bar + 16;          // per-thread staging of row_begin's inputs (EARLY_ROW == 2)
  static constexpr size_t bytes = off_row + P::ROWST_BYTES;
};

// Thread roles: threads 0..NT-1 are WORKERS (one quad of the strip each); the last warp is the AUXILIARY
// warp: it turns the quad totals of the published row into prefixes while the workers already add the
// next row to their running sums, and (plane readers) its lane 0 is the TMA producer of the row staging.
//
// One output row:   workers publish(yo) | bar A | aux scan(yo) || workers acc(row yin+1) | bar B |
//                   aux TMA(row yin+2) || workers window sums + per-pixel work of yo | bar C
template <class P>
__global__ void __launch_bounds__(P::NT + 32 * P::NAUX) __maxnreg__(P::MAXREG) gf_march_kernel(GfCommon gc, GfGeom gg) {
  constexpr int NI = P::NI, ND = P::ND, NT = P::NT, GP = GF_GP;
  constexpr int NIa = NI > 0 ? NI : 1;
  typedef GfSmem<P> L;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* Pd01 = reinterpret_cast<double2*>(smem_raw);
  double* Pd2 = reinterpret_cast<double*>(smem_raw + L::off_d23);
  double* Gd = reinterpret_cast<double*>(smem_raw + L::off_gd);
  uint4* Pi = reinterpret_cast<uint4*>(smem_raw + L::off_pi);
  uint32_t* Gi = reinterpret_cast<uint32_t*>(smem_raw + L::off_gi);
  auto select_buffer = [&](int yo) {  // published-row arrays of output row yo
    unsigned char* b = smem_raw + ((P::DBUF && (yo & 1)) ? L::buf_bytes : 0);
    Pd01 = reinterpret_cast<double2*>(b);
    Pd2 = reinterpret_cast<double*>(b + L::off_d23);
    Gd = reinterpret_cast<double*>(b + L::off_gd);
    Pi = reinterpret_cast<uint4*>(b + L::off_pi);
    Gi = reinterpret_cast<uint32_t*>(b + L::off_gi);
  };
  typename P::Shared* sh = reinterpret_cast<typename P::Shared*>(smem_raw + L::off_sh);
  unsigned char* stage = smem_raw + L::off_st;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L::off_bar);
  P pol;
  pol.init(gc, blockIdx.z, sh, gg);

  const int t = threadIdx.x;
  const bool aux = t >= NT;
  const int lane = t & 31;
  const int NQ = gg.NQ;
  const int W = gg.W, H = gg.H, r = gg.r;
  const int xs = blockIdx.x * gg.SW;
  const int gx = xs - gg.HL - 4 + 4 * t;      // image column of this thread's quad (multiple of 4)
  const bool qact = t < NQ;
  unsigned cmask = 0;                         // columns of the quad that are image columns; quad 0 is the zero guard
  if (qact && t > 0) {
#pragma unroll
    for (int c = 0; c < 4; c++) if (gx + c >= 0 && gx + c < W) cmask |= 1u << c;
  }
  const bool qload = cmask != 0;              // then 0 <= gx < Wp: the whole quad is readable
  const int ys = blockIdx.y * gg.seg_h, ye = min(ys + gg.seg_h, H);
  const int y_first = max(ys - r, 0);         // first row that enters the running sums
  const int y_begin = ys - r, y_end = ye + r; // rows yin of the march
  // output quads of this strip
  const int tq0 = gg.HL / 4 + 1;
  const bool oact = (t >= tq0) && (t < tq0 + gg.SW / 4) && (gx < W);
  // Everything the march loop needs to know about this thread's quad lives in ONE word (at 255 registers the compiler
  // otherwise parks the individual flags and counts in local-memory spill slots and reloads them every row):
  //   bits 28..31 cmask; bits 7c..7c+6 the horizontal pixel count of the window of column c, set for output quads only
  //   (so "output quad" == low 7 bits non-zero).  Counts need 2r + 1 <= 127; wider windows recompute them per row.
  const bool nx_packed = (2 * gg.r + 1) <= 127;
  unsigned meta = cmask << 28;
  if (oact) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      int x = gx + c;
      int nx = (x < gg.W) ? min(x + gg.r, gg.W - 1) - max(x - gg.r, 0) + 1 : 0;
      meta |= (unsigned)(nx_packed ? nx : 1) << (7 * c);
    }
  }
  const int rho = r >> 2;
  // addresses used by the window sums: the quads at -rho / +rho and the totals around them
  const int tlo_h = oact ? t - rho : 0, thi_h = oact ? t + rho : 0;  // policies without META keep them across the march
  // quads of this strip that lie inside the padded image: [tA, tB) (the guard quad 0 stays zero)
  const int tA = max(1, (gg.HL + 4 - xs) >> 2), tB = min(NQ, (gg.Wp - xs + gg.HL + 4) >> 2);

  // the quad-total rows are scanned over their whole length: keep the unused tail finite
  for (int bsel = 0; bsel < (P::DBUF ? 2 : 1); bsel++) {
    select_buffer(bsel);
    for (int i = t; i < GP; i += NT + 32 * P::NAUX) {
#pragma unroll
      for (int k = 0; k < NI; k++) Gi[k * GP + i] = 0u;
#pragma unroll
      for (int k = 0; k < ND; k++) Gd[k * GP + i] = 0.0;
    }
  }
  if constexpr (!P::PREFETCH) {
    if (t == NT) mbar_init(mbar, 1);
  }
  __syncthreads();

  uint32_t Vi[4][NIa];
  double Vd[4][ND];
#pragma unroll
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int k = 0; k < NIa; k++) Vi[c][k] = 0;
#pragma unroll
    for (int k = 0; k < ND; k++) Vd[c][k] = 0.0;
  }

  // TMA producer of the plane readers (aux lane 0): the entering and the leaving row of every plane for
  // march row `yi` land in the staging buffer and complete one mbarrier phase
  auto tma_rows = [&](int yi) {
    if constexpr (!P::PREFETCH) {
      if (yi >= y_end) return;
      const int yli = yi - 2 * r - 1;
      const bool en = (yi >= 0 && yi < H), le = (yli >= y_first);
      // one bulk copy per row: (tB - tA) quads x NP chunks of 16 bytes, contiguous in the interleaved layout
      const unsigned row_bytes = (unsigned)(tB - tA) * 16u * P::NP;
      mbar_arrive_expect_tx(mbar, ((en ? 1u : 0u) + (le ? 1u : 0u)) * row_bytes);
      float4* stg = reinterpret_cast<float4*>(stage);
      const int gqa = (xs - gg.HL - 4 + 4 * tA) >> 2;   // global quad index of strip quad tA
      const float4* rows = reinterpret_cast<const float4*>(pol.coef_rows());
      const size_t qpr = (size_t)(gg.Wp >> 2);
      if (en) tma_bulk_g2s(stg + (size_t)tA * P::NP, rows + ((size_t)yi * qpr + gqa) * P::NP, row_bytes, mbar);
      if (le) tma_bulk_g2s(stg + (size_t)(NT + tA) * P::NP, rows + ((size_t)yli * qpr + gqa) * P::NP, row_bytes, mbar);
    }
  };
  // workers: add march row yi to the running sums and drop row yi - 2r - 1
  auto acc = [&](int yi) {
    if (yi >= y_end) return;
    const int yli = yi - 2 * r - 1;
    const bool enter = (yi >= 0 && yi < H), leave = (yli >= y_first);
    const int par = (yi - y_begin) & 1;
    if constexpr (P::PREFETCH) {
      // this row was requested one march row ago (cp.async, no registers held meanwhile); the next one
      // goes out now into the other buffer
      cp_async_wait_all();
      typename P::Raw curE, curL;
      // UNCOND_STAGE: read unconditionally (a slot that was not filled for this row is read but never used) - GF1a's
      // conditionally assigned struct is otherwise kept in local memory by the compiler
      if (P::UNCOND_STAGE || (qload && enter)) pol.stage_read(stage, 2 * par, curE);
      if (P::UNCOND_STAGE || (qload && leave)) pol.stage_read(stage, 2 * par + 1, curL);
      const int yn = yi + 1, yln = yli + 1;
      if (qload && yn < y_end) {
        int gxa = gx;  // opaque: the row addresses are formed from the (uniform) plane base here, not carried per thread
        if constexpr (P::META) asm("" : "+r"(gxa) : "r"(yi));
        if (yn >= 0 && yn < H) pol.stage_issue(stage, 2 * (par ^ 1), yn, gxa);
        if (yln >= y_first) pol.stage_issue(stage, 2 * (par ^ 1) + 1, yln, gxa);
      }
      cp_async_commit();
      unsigned cm = meta;  // opaque copy: keeps the four bit tests of the partial-quad path out of spill slots
      if constexpr (P::META) asm("" : "+r"(cm) : "r"(yi));  // not volatile (free to schedule), tied to the row so that it stays in the loop
      cm >>= 28;
      if (cm == 0xfu) {  // whole quad inside the image: straight-line code
        if (enter) pol.template accum<+1, true>(curE, cm, Vi, Vd);
        if (leave) pol.template accum<-1, true>(curL, cm, Vi, Vd);
      } else if (qload) {
        if (enter) pol.template accum<+1, false>(curE, cm, Vi, Vd);
        if (leave) pol.template accum<-1, false>(curL, cm, Vi, Vd);
      }
    } else {
      mbar_wait(mbar, (unsigned)par);  // phase parity = march row parity
      if (qload && (enter || leave)) gf_accum_staged<P::NP, NT>(reinterpret_cast<const float4*>(stage), gx >> 2, enter, leave, cmask, Vd);
    }
  };

  // ---- prologue: first march row ------------------------------------------------------------------
  if constexpr (P::PREFETCH) {
    if (!aux) {
      int yl = y_begin - 2 * r - 1;
      if (qload && y_begin >= 0 && y_begin < H) pol.stage_issue(stage, 0, y_begin, gx);
      if (qload && yl >= y_first) pol.stage_issue(stage, 1, yl, gx);
      cp_async_commit();
      acc(y_begin);
    }
  } else {
    if (t == NT) tma_rows(y_begin);
    if (!aux) acc(y_begin);
    __syncthreads();
    if (t == NT) tma_rows(y_begin + 1);
  }

  // named barrier 1: the workers ARRIVE (they do not wait) once their row is published, the auxiliary warp waits on it
  auto bar_arrive_published = [&]() { asm volatile("bar.arrive 1, %0;" ::"n"(NT + 32 * P::NAUX) : "memory"); };
  auto bar_wait_published = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NT + 32 * P::NAUX) : "memory"); };
  // named barrier 2 (plane readers): the workers arrive when they have consumed the staged rows, the TMA producer waits
  auto bar_arrive_consumed = [&]() { asm volatile("bar.arrive 2, %0;" ::"n"(NT) : "memory"); };
  auto bar_wait_consumed = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); };
  static_assert(P::DBUF || !P::DELAY, "the delayed window phase reads one published row while the next one is written");

  // window sums and per-pixel work of output row yo (published and scanned during the previous loop iteration)
  auto window_phase = [&](const int yo) {
    select_buffer(yo);
    // ---- window sums and the per-pixel work -----------------------------------------------------------
    unsigned mrow = meta;  // opaque per-row copy (see the definition of meta)
    if constexpr (P::META) asm("" : "+r"(mrow) : "r"(yo));
    if (P::META ? (mrow & 127u) != 0 : oact) {  // output quad
      // META: the two shared-memory indices are formed from the thread index per row (held across the march they end up
      // in spill slots)
      int tq = t;
      if constexpr (P::META) asm("" : "+r"(tq) : "r"(yo));
      const int tlo = P::META ? tq - rho : tlo_h, thi = P::META ? tq + rho : thi_h;
      if constexpr (P::EARLY_ROW == 0) pol.row_begin(yo, gx);
      if constexpr (P::EARLY_ROW == 2) pol.row_pickup(smem_raw + L::off_row);
      // integer window sums: all four columns at once (two 16-byte loads per moment), or - for policies
      // that trade loads for registers (INT_HALF) - two columns per half
      uint32_t si[P::INT_HALF ? 2 : 4][NIa];
      auto int_sums = [&](int h) {
        if (gg.fast) {
#pragma unroll
          for (int k = 0; k < NI; k++) {
            uint32_t Wq = Gi[k * GP + thi - 1] - Gi[k * GP + tlo - 1];
            uint4 a = Pi[k * NT + tlo];
            if constexpr (!P::INT_HALF) {
              uint4 b = Pi[k * NT + thi];
              si[0][k] = Wq + b.x; si[1][k] = Wq - a.x + b.y; si[2][k] = Wq - a.y + b.z; si[3][k] = Wq - a.z + b.w;
            } else {
              uint2 b = reinterpret_cast<const uint2*>(Pi + k * NT + thi)[h];
              if (h == 0) { si[0][k] = Wq + b.x; si[1][k] = Wq - a.x + b.y; }
              else { si[0][k] = Wq - a.y + b.x; si[1][k] = Wq - a.z + b.y; }
            }
          }
        } else {
          const uint32_t* Pis = reinterpret_cast<const uint32_t*>(Pi);
#pragma unroll
          for (int cc = 0; cc < (P::INT_HALF ? 2 : 4); cc++) {
            int c = P::INT_HALF ? 2 * h + cc : cc;
            int zl = 4 * t + c - r, zh = 4 * t + c + r + 1;
#pragma unroll
            for (int k = 0; k < NI; k++) {
              uint32_t fl = Gi[k * GP + (zl >> 2) - 1] + ((zl & 3) ? Pis[k * NT * 4 + zl - 1] : 0u);
              uint32_t fh = Gi[k * GP + (zh >> 2) - 1] + ((zh & 3) ? Pis[k * NT * 4 + zh - 1] : 0u);
              si[cc][k] = fh - fl;
            }
          }
        }
      };
      if constexpr (!P::INT_HALF) int_sums(0);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if constexpr (P::INT_HALF) int_sums(h);
        double sd[2][ND];
        if (gg.fast) {
#pragma unroll
          for (int k = 0; k < ND; k++) {
            const double Gl = Gd[k * GP + tlo - 1];
            double Wq = Gd[k * GP + thi - 1] - Gl;
            double2 a01 = Pd01[k * NT + tlo];
            if (h == 0) {
              double2 b = Pd01[k * NT + thi];
              sd[0][k] = Wq + b.x; sd[1][k] = (Wq - a01.x) + b.y;
            } else {
              double a2 = Pd2[k * NT + tlo], b2 = Pd2[k * NT + thi];
              sd[0][k] = (Wq - a01.y) + b2;
              sd[1][k] = (Gd[k * GP + thi] - Gl) - a2;  // the whole quad thi is inside: totals up to and including it
            }
          }
        } else {
          const double* P01 = reinterpret_cast<const double*>(Pd01);
#pragma unroll
          for (int cc = 0; cc < 2; cc++) {
            int c = 2 * h + cc;
            int zl = 4 * t + c - r, zh = 4 * t + c + r + 1;
#pragma unroll
            for (int k = 0; k < ND; k++) {
              // prefix element (z&3)-1 of quad z>>2: elements 0,1 live in Pd01, 2 in Pd2
              int el = (zl & 3) - 1, eh = (zh & 3) - 1;
              double pl = (el < 0) ? 0.0 : (el < 2 ? P01[(k * NT + (zl >> 2)) * 2 + el] : Pd2[k * NT + (zl >> 2)]);
              double ph = (eh < 0) ? 0.0 : (eh < 2 ? P01[(k * NT + (zh >> 2)) * 2 + eh] : Pd2[k * NT + (zh >> 2)]);
              sd[cc][k] = (Gd[k * GP + (zh >> 2) - 1] + ph) - (Gd[k * GP + (zl >> 2) - 1] + pl);
            }
          }
        }
        const unsigned cm = mrow >> 28;
        int gxo = gx;
        if constexpr (P::META) asm("" : "+r"(gxo) : "r"(yo));
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
          const int c = 2 * h + cc;
          if (cm & (1u << c)) {  // column gx + c is an image column
            const int x = gxo + c;
            const int ny = min(yo + r, H - 1) - max(yo - r, 0) + 1;
            const int nx = nx_packed ? (int)((mrow >> (7 * c)) & 127u) : min(x + r, W - 1) - max(x - r, 0) + 1;
            pol.column(cc, yo, x, ny * nx, si[P::INT_HALF ? cc : 2 * h + cc], sd[cc]);
          }
        }
        if (cm & (1u << (2 * h))) pol.store_pair(yo, gxo + 2 * h);
      }
    }
  };

  if constexpr (P::DELAY) {
    // One loop iteration = one output row, ONE CTA-wide barrier:
    //   workers: publish(yo) | arrive 1 | add march row yin+1 | arrive 2 | window sums + per-pixel work of row yo-1 | barrier
    //   aux:     wait 1 (published) | scan the quad totals of row yo | wait 2 (staging consumed) | TMA(yin+2)       | barrier
    // The scan of row yo has the whole worker phase to finish and is consumed one iteration later, so the workers never
    // wait for it; the published rows are double-buffered by row parity, which is what makes the delay possible.
    // invariant at the top: the sums hold the march rows <= yin; row yin+1 has been requested
    for (int yin = y_begin; yin <= y_end; ++yin) {  // the extra iteration runs the window phase of the last row
      const int yo = yin - r;
      if (yo < ys) {  // warm-up rows (uniform across the CTA)
        if (!aux) acc(yin + 1);
        if constexpr (!P::PREFETCH) {
          __syncthreads();
          if (t == NT) tma_rows(yin + 2);
        }
        continue;
      }
      const bool pub = yo < ye;  // uniform
      if (pub) {
        // ---- publish the quad prefixes and totals -------------------------------------------------------
        select_buffer(yo);
        if (qact) {
  #pragma unroll
          for (int k = 0; k < NI; k++) {
            uint32_t p0 = Vi[0][k], p1 = p0 + Vi[1][k], p2 = p1 + Vi[2][k], p3 = p2 + Vi[3][k];
            Pi[k * NT + t] = make_uint4(p0, p1, p2, p3);
            Gi[k * GP + t] = p3;
          }
  #pragma unroll
          for (int k = 0; k < ND; k++) {
            double p0 = Vd[0][k], p1 = p0 + Vd[1][k], p2 = p1 + Vd[2][k], p3 = p2 + Vd[3][k];
            Pd01[k * NT + t] = make_double2(p0, p1);
            Pd2[k * NT + t] = p2;
            Gd[k * GP + t] = p3;
          }
        }
      }
      if (aux) {
        if (pub) {
          bar_wait_published();
          // prefix over the quad totals, four moments per round (8 lanes each); the rounds alternate between
          // the auxiliary warps
          const int seg = lane & 7, mq = lane >> 3, aw = (t - NT) >> 5;
          constexpr int RI = (NI + 3) / 4, RD = (ND + 3) / 4;
  #pragma unroll
          for (int rd = 0; rd < RI + RD; rd++) {
            if (rd % P::NAUX != aw) continue;
            if (rd < RI) { const int k0 = 4 * rd; gf_scan_task<uint32_t, uint4>(Gi + min(k0 + mq, NIa - 1) * GP, seg, k0 + mq < NI); }
            else { const int k0 = 4 * (rd - RI); gf_scan_task<double, double2>(Gd + min(k0 + mq, ND - 1) * GP, seg, k0 + mq < ND); }
          }
        }
      } else {
        if (pub) bar_arrive_published();
        acc(yin + 1);
        if constexpr (!P::PREFETCH) {
          // plane readers: the LAST worker warp waits until every worker has consumed the staged rows (named barrier 2,
          // workers only) and its first lane requests the next ones, which land during the window phase
          if (t >= NT - 32) {
            bar_wait_consumed();
            if (t == NT - 32) tma_rows(yin + 2);
          } else {
            bar_arrive_consumed();
          }
        }
        if (yo > ys) window_phase(yo - 1);
      }
      __syncthreads();
    }
  } else {
    // invariant at the top: the sums hold the march rows <= yin; row yin+1 has been requested
    for (int yin = y_begin; yin < y_end; ++yin) {
      const int yo = yin - r;
      if (yo < ys) {  // warm-up rows (uniform across the CTA)
        if (!aux) acc(yin + 1);
        if constexpr (!P::PREFETCH) {
          __syncthreads();
          if (t == NT) tma_rows(yin + 2);
        }
        continue;
      }
      // ---- publish the quad prefixes and totals -------------------------------------------------------
      select_buffer(yo);
      if constexpr (P::EARLY_SCAN) {
        // quad totals first: as soon as every worker has stored them (named barrier 1, the workers only arrive) the
        // auxiliary warp starts its scan, while the workers go on with the quad prefixes and the next march row.
        // The sums are exact, so (V0 + V2) + (V1 + V3) is the last element of the prefix chain bit for bit (and shares no
        // subexpression with it that the compiler would keep in registers across the hand-off).
        if (qact) {
  #pragma unroll
          for (int k = 0; k < NI; k++) Gi[k * GP + t] = (Vi[0][k] + Vi[2][k]) + (Vi[1][k] + Vi[3][k]);
  #pragma unroll
          for (int k = 0; k < ND; k++) Gd[k * GP + t] = (Vd[0][k] + Vd[2][k]) + (Vd[1][k] + Vd[3][k]);
        }
        if (aux) bar_wait_published(); else bar_arrive_published();
        if (qact) {
  #pragma unroll
          for (int k = 0; k < NI; k++) {
            uint32_t p0 = Vi[0][k], p1 = p0 + Vi[1][k], p2 = p1 + Vi[2][k], p3 = p2 + Vi[3][k];
            Pi[k * NT + t] = make_uint4(p0, p1, p2, p3);
          }
  #pragma unroll
          for (int k = 0; k < ND; k++) {
            double p0 = Vd[0][k], p1 = p0 + Vd[1][k], p2 = p1 + Vd[2][k];
            Pd01[k * NT + t] = make_double2(p0, p1);
            Pd2[k * NT + t] = p2;
          }
        }
      } else {
        if (qact) {
    #pragma unroll
          for (int k = 0; k < NI; k++) {
            uint32_t p0 = Vi[0][k], p1 = p0 + Vi[1][k], p2 = p1 + Vi[2][k], p3 = p2 + Vi[3][k];
            Pi[k * NT + t] = make_uint4(p0, p1, p2, p3);
            Gi[k * GP + t] = p3;
          }
    #pragma unroll
          for (int k = 0; k < ND; k++) {
            double p0 = Vd[0][k], p1 = p0 + Vd[1][k], p2 = p1 + Vd[2][k], p3 = p2 + Vd[3][k];
            Pd01[k * NT + t] = make_double2(p0, p1);
            Pd2[k * NT + t] = p2;
            Gd[k * GP + t] = p3;
          }
        }
        __syncthreads();  // A
      }
      if (aux) {
        // prefix over the quad totals, four moments per round (8 lanes each); the rounds alternate between
        // the auxiliary warps
        const int seg = lane & 7, mq = lane >> 3, aw = (t - NT) >> 5;
        constexpr int RI = (NI + 3) / 4, RD = (ND + 3) / 4;
  #pragma unroll
        for (int rd = 0; rd < RI + RD; rd++) {
          if (rd % P::NAUX != aw) continue;
          if (rd < RI) { const int k0 = 4 * rd; gf_scan_task<uint32_t, uint4>(Gi + min(k0 + mq, NIa - 1) * GP, seg, k0 + mq < NI); }
          else { const int k0 = 4 * (rd - RI); gf_scan_task<double, double2>(Gd + min(k0 + mq, ND - 1) * GP, seg, k0 + mq < ND); }
        }
      } else {
        // plane readers that want it: the global loads of row yo's per-pixel inputs go out here, a whole accumulate phase
        // (and barrier B) before their first use
        if constexpr (P::EARLY_ROW == 1) { if (oact) pol.row_begin(yo, gx); }
        if constexpr (P::EARLY_ROW == 2) { if (oact) pol.row_prefetch(smem_raw + L::off_row, yo, gx); }
        acc(yin + 1);
      }
      __syncthreads();  // B
      if constexpr (!P::PREFETCH) {
        if (t == NT) tma_rows(yin + 2);  // every worker has consumed the staged rows
      }
      window_phase(yo);
      if constexpr (!P::DBUF) __syncthreads();  // C (with two buffers the barriers A, B of the next row order the reuse)
    }
  }
  pol.finish();
}

// geometry + launch -------------------------------------------------------------------------------------
static GfGeom gf_geometry(int W, int H, int r, int NT) {
  GfGeom g;
  g.W = W; g.H = H; g.Wp = (W + 3) & ~3; g.r = r;
  g.HL = (r + 3) & ~3;
  int sw_max = 4 * (NT - 1) - 2 * g.HL;
  int strips = cdiv(W, sw_max);
  g.SW = (cdiv(W, strips) + 3) & ~3;
  g.NQ = 1 + (2 * g.HL + g.SW) / 4;
  g.seg_h = H;
  g.fast = (r % 4 == 0) ? 1 : 0;
  return g;
}

static_assert(GfSmem<PolGF2b>::bytes + 1024 <= (227 * 1024) / 2, "GF2b runs two CTAs per SM");
static_assert(GfSmem<PolGF1a>::bytes <= 227 * 1024 && GfSmem<PolGF1b>::bytes <= 227 * 1024 && GfSmem<PolGF2a>::bytes <= 227 * 1024, "shared memory");

template <class P>
static int gf_launch(uwip_ctx* ctx, const char* tag, const GfCommon& gc, int n, int W, int H, int r) {
  static_assert(P::NT <= GF_GP, "quad-total rows hold one entry per thread");
  GfGeom gg = gf_geometry(W, H, r, P::NT);
  int strips = cdiv(W, gg.SW);
  // vertical segments: fill the machine (tail of the last wave) against the 2r warm-up rows per segment
  int slots = ctx->sm_count * P::MINB;
  int tasks = strips * n;
  int best = 1;
  double best_eff = 0.0;
  int max_segs = std::max(1, H / (4 * r + 2));
  for (int s = 1; s <= max_segs && s <= 64; s++) {
    int sh = cdiv(H, s), sc = cdiv(H, sh);
    double waves = (double)cdiv(tasks * sc, slots);
    double eff = ((double)tasks * sc / (waves * slots)) * ((double)sh / (double)(sh + 2 * r));
    if (eff > best_eff * 1.02) { best_eff = eff; best = s; }
  }
  gg.seg_h = cdiv(H, best);
  int segs = cdiv(H, gg.seg_h);
  size_t smem = GfSmem<P>::bytes;
  static bool attr_done = false;  // per template instantiation
  if (!attr_done) {
    UWIP_CUDA(ctx, cudaFuncSetAttribute(gf_march_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid(strips, segs, n);
  UWIP_LAUNCH(ctx, tag, gf_march_kernel<P>, grid, P::NT + 32 * P::NAUX, smem, gc, gg);
  return UWIP_OK;
}

// -------------------------------------------------------------------------------------------------
// E: restored -> R8, I8 -> YCrCb joint min / max (BGDehaze.py:75-80); stores (Yi, Cri, Cbi, Yj)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) exposure_minmax_kernel(GfCommon g, int W, int H, int Wp, double* dbg_restored) {
  __shared__ ExpShared sh;
  int f = blockIdx.y;
  size_t n_pp = (size_t)Wp * H;
  exp_shared_init(&sh, g.fs[f], (double)W * (double)H);
  const uint32_t* kq = g.kq + (size_t)f * n_pp;
  const float* J = g.J + (size_t)f * 2 * n_pp;
  uint32_t* ycc = g.ycc + (size_t)f * n_pp;
  unsigned imn = 255, imx = 0, jmn = 255, jmx = 0;
  const int qpr = Wp / 4;
  const size_t n_q = (size_t)qpr * H;
  for (size_t qi = (size_t)blockIdx.x * 256 + threadIdx.x; qi < n_q; qi += (size_t)gridDim.x * 256) {
    int y = (int)(qi / qpr), x0 = (int)(qi - (size_t)y * qpr) * 4;
    size_t pp = (size_t)y * Wp + x0;
    uint4 kw = __ldg(reinterpret_cast<const uint4*>(kq + pp));
    float4 jb = __ldg(reinterpret_cast<const float4*>(J + pp)), jg = __ldg(reinterpret_cast<const float4*>(J + n_pp + pp));
    uint32_t out[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      out[c] = 0;
      if (x0 + c >= W) continue;
      uint32_t w = quad_get(kw, c);
      uint32_t kb = w & 255u, kg = (w >> 8) & 255u, kr = (w >> 16) & 255u;
      int r8b = restored_byte(quad_get(jb, c), sh.fc, 0), r8g = restored_byte(quad_get(jg, c), sh.fc, 1), r8r = sh.rt.red8[kr];
      if (dbg_restored && f == 0) {
        size_t pix = (size_t)y * W + x0 + c;
        dbg_restored[pix * 3] = norm_j(quad_get(jb, c), sh.fc, 0); dbg_restored[pix * 3 + 1] = norm_j(quad_get(jg, c), sh.fc, 1);
        dbg_restored[pix * 3 + 2] = sh.rt.redN[kr];
      }
      int Yi, Cri, Cbi, Yj, Crj, Cbj;
      bgr2ycrcb_u8(sh.rt.i8[kb], sh.rt.i8[kg], sh.rt.i8[kr], Yi, Cri, Cbi);
      imn = min(imn, (unsigned)imin3(Yi, Cri, Cbi)); imx = max(imx, (unsigned)imax3(Yi, Cri, Cbi));
      bgr2ycrcb_u8(r8b, r8g, r8r, Yj, Crj, Cbj);
      jmn = min(jmn, (unsigned)imin3(Yj, Crj, Cbj)); jmx = max(jmx, (unsigned)imax3(Yj, Crj, Cbj));
      out[c] = (uint32_t)Yi | ((uint32_t)Cri << 8) | ((uint32_t)Cbi << 16) | ((uint32_t)Yj << 24);
    }
    *reinterpret_cast<uint4*>(ycc + pp) = make_uint4(out[0], out[1], out[2], out[3]);
  }
  imn = warp_reduce_min_u32(imn); imx = warp_reduce_max_u32(imx);
  jmn = warp_reduce_min_u32(jmn); jmx = warp_reduce_max_u32(jmx);
  if ((threadIdx.x & 31) == 0) {
    FrameState& s = g.fs[f];
    atomicMin(&s.yi_min, imn); atomicMax(&s.yi_max, imx);
    atomicMin(&s.yj_min, jmn); atomicMax(&s.yj_max, jmx);
  }
}

// S (BGDehaze.py:83) as a table over (Yi - min, Yj - min): both are bytes.  grid (256, n), block 256.
__global__ void __launch_bounds__(256) stab_kernel(const FrameState* __restrict__ fs, double* __restrict__ stab) {
  int f = blockIdx.y, gy = blockIdx.x, j = threadIdx.x;
  const FrameState& s = fs[f];
  double yi_rng = (double)((int)s.yi_max - (int)s.yi_min), yj_rng = (double)((int)s.yj_max - (int)s.yj_min);
  double yi = (double)gy / yi_rng, yj = (double)j / yj_rng;
  double yi2 = 0.3 * (yi * yi);
  stab[(size_t)f * 65536 + gy * 256 + j] = (yj * yi + yi2) / (yj * yj + yi2);
}

// S per pixel as an f32 plane: GF2a then reads it through the same staged row copies as the guide instead
// of gathering from the table inside the march (f32 keeps 2^-24 relative: far inside the 1e-5 budget).
__global__ void __launch_bounds__(256) splane_kernel(GfCommon g, int H, int Wp) {
  int f = blockIdx.y;
  size_t n_pp = (size_t)Wp * H;
  const FrameState& s = g.fs[f];
  const uint32_t a = s.yi_min, b = s.yj_min;
  const uint32_t ysub = a | (a << 8) | (a << 16) | (b << 24);
  const uint32_t* ycc = g.ycc + (size_t)f * n_pp;
  const double* stab = g.stab + (size_t)f * 65536;
  float* sp = g.splane + (size_t)f * n_pp;
  const size_t n_q = n_pp / 4;
  for (size_t qi = (size_t)blockIdx.x * 256 + threadIdx.x; qi < n_q; qi += (size_t)gridDim.x * 256) {
    uint4 yw = __ldg(reinterpret_cast<const uint4*>(ycc) + qi);
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t w = quad_get(yw, c);
      // pad columns hold zeros: keep them away from the table (their S is never used)
      bool pad = (w & 255u) < a || (w >> 24) < b;
      w -= ysub;
      o[c] = pad ? 0.f : (float)grid_round<28>(__ldg(stab + (((w & 255u) << 8) | (w >> 24))));
    }
    reinterpret_cast<float4*>(sp)[qi] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// final: (OutputExp - min)/(max - min) * 255 -> rint -> saturate (BGDehaze.py:88-89, main.py:19)
__global__ void __launch_bounds__(256, 4) final_kernel(GfCommon g, int W, int H, int Wp, uint8_t* __restrict__ dst, double* dbg_out, int32_t* __restrict__ flags) {
  __shared__ ExpShared sh;
  int f = blockIdx.y;
  size_t n_pp = (size_t)Wp * H;
  exp_shared_init(&sh, g.fs[f], (double)W * (double)H);
  const FrameState& s = g.fs[f];
  double omn = dunkey(s.omin_key), den = dunkey(s.omax_key) - omn;
  double scale = 255.0 / den;
  bool nan_frame = s.nan_flag != 0;
  if (flags && blockIdx.x == 0 && threadIdx.x == 0) flags[f] = nan_frame ? UWIP_FRAME_NAN : 0;
  const uint32_t* kq = g.kq + (size_t)f * n_pp;
  const float* J = g.J + (size_t)f * 2 * n_pp;
  const float* refS = g.refS + (size_t)f * n_pp;
  uint8_t* out = dst + (size_t)f * W * H * 3;
  const int qpr = Wp / 4;
  const size_t n_q = (size_t)qpr * H;
  const bool vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
  for (size_t qi = (size_t)blockIdx.x * 256 + threadIdx.x; qi < n_q; qi += (size_t)gridDim.x * 256) {
    int y = (int)(qi / qpr), x0 = (int)(qi - (size_t)y * qpr) * 4;
    size_t pp = (size_t)y * Wp + x0;
    uint4 kw = __ldg(reinterpret_cast<const uint4*>(kq + pp));
    float4 jb = __ldg(reinterpret_cast<const float4*>(J + pp)), jg = __ldg(reinterpret_cast<const float4*>(J + n_pp + pp));
    float4 qs = __ldg(reinterpret_cast<const float4*>(refS + pp));
    uint8_t b[12];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double rest[3];
      rest[0] = norm_j_fast(quad_get(jb, c), sh.fc, 0);
      rest[1] = norm_j_fast(quad_get(jg, c), sh.fc, 1);
      rest[2] = sh.rt.redN[(quad_get(kw, c) >> 16) & 255u];
      double q = (double)quad_get(qs, c);
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        double v = (rest[ch] * q - omn) * scale;
        if (dbg_out && f == 0 && x0 + c < W) dbg_out[((size_t)y * W + x0 + c) * 3 + ch] = nan_frame ? __longlong_as_double(0x7ff8000000000000ll) : (rest[ch] * q - omn) / den;
        int bv = 0;
        if (!nan_frame && v == v && fabs(v) < 2.0e9) bv = min(max(__double2int_rn(v), 0), 255);
        b[c * 3 + ch] = (uint8_t)bv;
      }
    }
    uint8_t* o = out + ((size_t)y * W + x0) * 3;
    if (vec_ok) {
      uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
      o32[0] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
      o32[1] = (uint32_t)b[4] | ((uint32_t)b[5] << 8) | ((uint32_t)b[6] << 16) | ((uint32_t)b[7] << 24);
      o32[2] = (uint32_t)b[8] | ((uint32_t)b[9] << 8) | ((uint32_t)b[10] << 16) | ((uint32_t)b[11] << 24);
    } else {
      for (int c = 0; c < 4 && x0 + c < W; c++) { o[c * 3] = b[c * 3]; o[c * 3 + 1] = b[c * 3 + 1]; o[c * 3 + 2] = b[c * 3 + 2]; }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// driver
// -------------------------------------------------------------------------------------------------
// Batch size for which the strips of the one-CTA-per-SM marches make one full wave (4K: 148 SMs / 5 strips =
// 29 frames); the host-buffer pipeline cuts its sub-batches in multiples of it.
int dehaze_wave_frames(const uwip_ctx* ctx, int w) {
  GfGeom g = gf_geometry(w, 4 * 40 + 2, 40, PolGF1a::NT);
  return std::max(1, ctx->sm_count / cdiv(w, g.SW));
}

int dehaze_frames_dev(uwip_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, int n, int W, int H, const uwip_dehaze_params& p,
                      bool minmax_done, FrameState* fs, DehazeDebug* dbg, int32_t* d_flags) {
  UWIP_REQUIRE(ctx, n >= 1 && W >= 1 && H >= 1, "bad size");
  UWIP_REQUIRE(ctx, p.window >= 1 && p.window <= WK_MAXWIN, "window must be 1..33");
  UWIP_REQUIRE(ctx, p.radius >= 1 && p.radius <= 160, "radius must be 1..160");
  UWIP_REQUIRE(ctx, (size_t)W * H < (1ull << 31), "frame too large");
  UWIP_REQUIRE(ctx, !dbg || n == 1, "stage outputs are single-frame");
  size_t n_px = (size_t)W * H;
  const int Wp = (W + 3) & ~3;
  size_t n_pp = (size_t)Wp * H;
  if (!minmax_done) {
    int want = (int)std::min<size_t>((n_px * 3 / 16 + 255) / 256, 1u << 16);
    int gxm = std::max(1, std::min(want, ctx->sm_count * 8 / n + 1));
    dim3 grid(gxm, n);
    UWIP_LAUNCH(ctx, "dz_minmax", minmax_kernel, grid, 256, 0, d_src, n_px * 3, fs);
  }
  uint32_t* d_kq = (uint32_t*)uwip_slot(ctx, SLOT_KQ, (size_t)n * n_pp * 4);
  uint8_t* d_mg = (uint8_t*)uwip_slot(ctx, SLOT_MPLANES, (size_t)n * n_pp);
  uint32_t* d_ycc = (uint32_t*)uwip_slot(ctx, SLOT_YCC, (size_t)n * n_pp * 4);
  double* d_stab = (double*)uwip_slot(ctx, SLOT_STAB, (size_t)n * 65536 * 8);
  float* d_sp = (float*)uwip_slot(ctx, SLOT_SPLANE, (size_t)n * n_pp * sizeof(float));
  dim3 gridw(cdiv(Wp, WK_TX), cdiv(H, WK_TY), n);    // generic-window kernel
  dim3 gridf(cdiv(Wp, WF_TX), cdiv(H, WF_TY), n);    // 15x15 kernel
  int n_part = gridw.x * gridw.y, n_partf = gridf.x * gridf.y;
  ArgPartial* d_part = (ArgPartial*)uwip_slot(ctx, SLOT_PARTIALS, (size_t)n * std::max(n_part, n_partf) * sizeof(ArgPartial));
  float* d_ab = (float*)uwip_slot(ctx, SLOT_AB, (size_t)n * 8 * n_pp * sizeof(float));
  float* d_J = (float*)uwip_slot(ctx, SLOT_J, (size_t)n * 2 * n_pp * sizeof(float));
  float* d_refS = (float*)uwip_slot(ctx, SLOT_REFS, (size_t)n * n_pp * sizeof(float));
  if (!d_kq || !d_mg || !d_ycc || !d_stab || !d_sp || !d_part || !d_ab || !d_J || !d_refS) return UWIP_ERR_NOMEM;
  {
    const int wmax = p.window, wmin = WK_TWIN;
    if (wmax != WK_TWIN) {
      // Background_light(normI, w) for dehazed_BG's B (BGDehaze.py:51) uses the caller's window ...
      int PL = std::max(wmax / 2, wmin / 2), PR = std::max(wmax - 1 - wmax / 2, wmin - 1 - wmin / 2);
      int RW = WK_TX + PL + PR, RH = WK_TY + PL + PR;
      size_t smem = ((size_t)2 * RW * RH + (size_t)3 * RH * WK_TX) * 4;
      static size_t attr = 0;
      if (smem > attr) {
        UWIP_CUDA(ctx, cudaFuncSetAttribute(window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
      }
      UWIP_LAUNCH(ctx, "dz_window", window_kernel, gridw, WK_THREADS, smem, d_src, W, H, Wp, wmax, fs, d_kq, d_mg, d_part);
      UWIP_LAUNCH(ctx, "dz_bglight", bglight_finish_kernel, n, 256, 0, d_src, W, H, d_part, n_part, fs, 1, 0);
    }
    // ... but refined_t() is always called without w (BGDehaze.py:52): the transmission and the
    // background light inside transmission_map use the 15x15 window
    size_t smemf = ((size_t)2 * WF_RH * WF_PP + (size_t)3 * WF_RH * WF_HP) * 4;
    static bool attrf = false;
    if (!attrf) {
      UWIP_CUDA(ctx, cudaFuncSetAttribute(window15_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemf));
      attrf = true;
    }
    UWIP_LAUNCH(ctx, "dz_window", window15_kernel, gridf, 256, smemf, d_src, W, H, Wp, fs, d_kq, d_mg, d_part);
    UWIP_LAUNCH(ctx, "dz_bglight", bglight_finish_kernel, n, 256, 0, d_src, W, H, d_part, n_partf, fs, wmax == WK_TWIN ? 1 : 0, 1);
  }
  if (dbg && dbg->stop_after == 1) return UWIP_OK;
  if (dbg && dbg->t_raw) {
    dim3 g2(cdiv(W, 256), H);
    UWIP_LAUNCH(ctx, "dz_traw", traw_kernel, g2, 256, 0, d_kq, d_mg, W, H, Wp, fs, dbg->t_raw);
  }
  if (dbg && dbg->stop_after == 2) return UWIP_OK;

  GfCommon gc;
  gc.kq = d_kq; gc.mg = d_mg; gc.ycc = d_ycc; gc.stab = d_stab; gc.splane = d_sp; gc.ab = d_ab; gc.J = d_J; gc.refS = d_refS; gc.fs = fs;
  gc.eps = p.eps; gc.tmin = p.tmin; gc.dbg_tref = dbg ? dbg->t_ref : nullptr;
  UWIP_CHECK(gf_launch<PolGF1a>(ctx, "dz_gf1a", gc, n, W, H, p.radius));
  UWIP_CHECK(gf_launch<PolGF1b>(ctx, "dz_gf1b", gc, n, W, H, p.radius));
  if (dbg && dbg->stop_after == 3) return UWIP_OK;
  int gx = std::max(1, std::min((int)((n_pp / 4 + 1023) / 1024), std::max(1, ctx->sm_count * 24 / n)));
  dim3 grid_e(gx, n);
  UWIP_LAUNCH(ctx, "dz_exposure_minmax", exposure_minmax_kernel, grid_e, 256, 0, gc, W, H, Wp, dbg ? dbg->restored : (double*)nullptr);
  if (dbg && dbg->stop_after == 4) return UWIP_OK;
  dim3 grid_s(256, n);
  UWIP_LAUNCH(ctx, "dz_stab", stab_kernel, grid_s, 256, 0, fs, d_stab);
  UWIP_LAUNCH(ctx, "dz_splane", splane_kernel, grid_e, 256, 0, gc, H, Wp);
  UWIP_CHECK(gf_launch<PolGF2a>(ctx, "dz_gf2a", gc, n, W, H, p.radius));
  UWIP_CHECK(gf_launch<PolGF2b>(ctx, "dz_gf2b", gc, n, W, H, p.radius));
  UWIP_LAUNCH(ctx, "dz_final", final_kernel, grid_e, 256, 0, gc, W, H, Wp, d_dst, dbg ? dbg->out : (double*)nullptr, d_flags);
  return UWIP_OK;
}

// dehaze.cu - bgdehaze (Li et al. ICASSP-16 blue-green dehazing + red-channel correction + adaptive
// exposure map) for sm_100a.
//   reference: modules/bgdehaze/BGDehaze.py:14-89, modules/bgdehaze/guidedfilter.py:23-103,
//              modules/bgdehaze/main.py:16-19.  Stage names D0..D10 follow SURVEY.md 8a.
//
// Data flow per frame (all frame-global reductions land in FrameState, no host round trips):
//   minmax (D0)  ->  window max / arg-min partials / window-min planes (D1,D2)  ->  background light
//   -> GF1a: strip-march box sums of 17 moments + per-pixel 3x3 solve -> a,b planes (f32 x8)
//   -> GF1b: box(a,b) -> refined t (D3,D5) -> J (D6) + min/max/sum reductions -> J planes (f32 x2)
//   -> E: restored -> R8/I8 -> YCrCb joint min/max (D7,D8 first half)
//   -> GF2a: S map + 13 moments + solve -> a,b (f32 x4)   -> GF2b: box(a,b) -> refined S, exposure min/max
//   -> final: normalise, x255, rint, saturate -> bgr8 (D8 second half, D10)
//
// The guided filter's (2r+1)^2 box sums are computed by a "strip march": a CTA owns a strip of image
// columns (one thread per column, halo included), walks down the rows keeping the vertical running
// sums in registers (add the entering row, subtract the leaving row), and for every output row turns
// them into horizontal window sums through a two-level prefix scan in shared memory.  Guide moments
// are exact 32-bit integers (the guide is k/range with k uint8); everything involving the filtered
// signal accumulates in fp64.  The per-pixel solve works in "k units" (guide not divided by range)
// with eps_k = eps*range^2, on the exact integer numerators N*S_ij - S_i*S_j.
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
// D1 / D2: 15x15 window max (3 channels) -> arg-min partials; window min (blue, green) -> m planes
// ------------------------------------------------------------------------------------------------
constexpr int WK_TX = 64, WK_TY = 32, WK_THREADS = 256;
constexpr int WK_MAXWIN = 33;
constexpr int WK_TWIN = 15;  // transmission window: always 15 (BGDehaze.py:52 drops w)

struct ArgPartial {
  double d0, d1;
  unsigned int i0, i1;
};

__device__ __forceinline__ bool lex_less(double a, unsigned ia, double b, unsigned ib) { return (a < b) || (a == b && ia < ib); }

__global__ void __launch_bounds__(WK_THREADS) window_kernel(const uint8_t* __restrict__ src, int W, int H, int wmax,
                                                            const FrameState* __restrict__ fs, uint8_t* __restrict__ mplanes,
                                                            ArgPartial* __restrict__ partials) {
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
  for (int j = 0; j < WK_TY / (WK_THREADS / WK_TX); j++) {
    int yy = rg * (WK_TY / (WK_THREADS / WK_TX)) + j;
    int gy = y0 + yy;
    if (gx >= W || gy >= H) continue;
    uint32_t m0 = 0, m1 = 0, n0 = 0xffffffffu;
    for (int k = -pmax; k < wmax - pmax; k++) {
      m0 = __vmaxu2(m0, HX0[(yy + PL + k) * WK_TX + x]);
      m1 = max(m1, HX1[(yy + PL + k) * WK_TX + x]);
    }
    for (int k = -pmin; k < wmin - pmin; k++) n0 = __vminu2(n0, HN0[(yy + PL + k) * WK_TX + x]);
    size_t pix = (size_t)gy * W + gx;
    uint8_t* mp = mplanes + (size_t)f * 2 * W * H;
    mp[pix] = (uint8_t)(n0 & 0xffffu);
    mp[(size_t)W * H + pix] = (uint8_t)(n0 >> 16);
    // D (BGDehaze.py:20-21): max_R - max_B, max_R - max_G on the normalised image, in fp64
    double nr = s_nrm[(int)m1 - kmin];
    double d0 = nr - s_nrm[(int)(m0 & 0xffffu) - kmin];
    double d1 = nr - s_nrm[(int)(m0 >> 16) - kmin];
    unsigned idx = (unsigned)pix;
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
__global__ void traw_kernel(const uint8_t* __restrict__ mplanes, int W, int H, const FrameState* __restrict__ fs,
                            double* __restrict__ t_raw) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const FrameState& s = fs[0];
  int kmin = s.kmin;
  double range = (double)((int)s.kmax - kmin);
  const int pw = WK_TWIN / 2;
  bool touches = (x < pw) || (y < pw) || (x - pw + WK_TWIN - 1 >= W) || (y - pw + WK_TWIN - 1 >= H);
  size_t pix = (size_t)y * W + x;
  for (int c = 0; c < 2; c++) {
    int m = touches ? 0 : (int)mplanes[(size_t)c * W * H + pix] - kmin;
    t_raw[(size_t)c * W * H + pix] = 1.0 - ((double)m / range) / s.Bt[c];
  }
}

// ------------------------------------------------------------------------------------------------
// strip-march box-sum engine
// ------------------------------------------------------------------------------------------------
constexpr int GF_NT = 384;                              // threads = strip columns incl. 2r halo
constexpr int GF_SEG = 24;                              // columns per scan segment
constexpr int GF_NSEG = GF_NT / GF_SEG;                 // 16 segments
constexpr int GF_PITCH = GF_SEG * (GF_NSEG + 1);        // column c lives at (c % SEG) * 17 + c / SEG
static_assert(GF_NSEG == 16, "phase-1 warp layout assumes 16 segments");

// 3x3 symmetric solve shared by GF1a / GF2a.  Inputs are window sums in k units:
//   S[3] (sum k_i), SS[6] (sum k_i k_j: 00 01 02 11 12 22), N (window pixel count),
//   Sp (sum p), Sip[3] (sum k_i p).  Output a[3] (k units) and b.
__device__ __forceinline__ void gf_build_M(const uint32_t* si, double N, double epsN2, double* M, double* Sd) {
  Sd[0] = (double)si[0]; Sd[1] = (double)si[1]; Sd[2] = (double)si[2];
  // exact: N*S_ij and S_i*S_j are integers < 2^53
  M[0] = fma(N, (double)si[3], -Sd[0] * Sd[0]) + epsN2;
  M[1] = fma(N, (double)si[4], -Sd[0] * Sd[1]);
  M[2] = fma(N, (double)si[5], -Sd[0] * Sd[2]);
  M[3] = fma(N, (double)si[6], -Sd[1] * Sd[1]) + epsN2;
  M[4] = fma(N, (double)si[7], -Sd[1] * Sd[2]);
  M[5] = fma(N, (double)si[8], -Sd[2] * Sd[2]) + epsN2;
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
  rdet = 1.0 / det;
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

struct GfCommon {
  const uint8_t* src;     // bgr8 frames
  const uint8_t* mplanes; // [n][2][H*W]
  float* ab;              // [n][8][H*W]
  float* J;               // [n][2][H*W]
  float* refS;            // [n][H*W]
  FrameState* fs;
  double eps, tmin;
  double* dbg_tref;       // optional [2][H*W] (frame 0 only)
};

// ---- shared per-CTA frame constants --------------------------------------------------------------
struct FrameConst {
  int kmin, range;
  double B[3], Bt[3];
  // restored-image parameters (valid after GF1b): J min / 1/(max-min), red LUT scalars
  double jmin[2], jinv[2];
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

// restored blue / green from the stored J value (dehazed_BG, BGDehaze.py:53-56)
__device__ __forceinline__ double norm_j(float j, const FrameConst& fc, int c) { return ((double)j - fc.jmin[c]) / fc.jinv[c]; }

// -------------------------------------------------------------------------------------------------
// policies
// -------------------------------------------------------------------------------------------------
// GF1a: guide = normI (k units), p = max(t_blue, tmin) and max(t_green, tmin)
struct PolGF1a {
  static constexpr int NI = 9, ND = 8;
  struct Shared {
    double pT[2][256];  // p_c as a function of the window-min k'
    FrameConst fc;
  };
  GfCommon g; Shared* sh; int W, H, f;
  const uint8_t* img; const uint8_t* mp; float* ab;
  double epsN_k;  // eps * range^2
  __device__ void init(const GfCommon& gc, int frame, Shared* s, int W_, int H_) {
    g = gc; sh = s; W = W_; H = H_; f = frame;
    size_t n_px = (size_t)W * H;
    img = g.src + (size_t)f * n_px * 3;
    mp = g.mplanes + (size_t)f * 2 * n_px;
    ab = g.ab + (size_t)f * 8 * n_px;
    if (threadIdx.x == 0) load_frame_const(g.fs[f], sh->fc);
    __syncthreads();
    double range = (double)sh->fc.range;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      int c = i >> 8, k = i & 255;
      double t = 1.0 - ((double)k / range) / sh->fc.Bt[c];     // transmission_map (BGDehaze.py:35-36)
      sh->pT[c][k] = (t < g.tmin) ? g.tmin : t;              // np.maximum(t, tmin) (NaN stays NaN)
    }
    epsN_k = g.eps * range * range;
    __syncthreads();
  }
  template <int SIGN>
  __device__ __forceinline__ void accum(int y, int x, uint32_t* Vi, double* Vd) const {
    size_t pix = (size_t)y * W + x;
    const uint8_t* p = img + pix * 3;
    int kmin = sh->fc.kmin;
    int kb = (int)p[0] - kmin, kg = (int)p[1] - kmin, kr = (int)p[2] - kmin;
    const int pw = WK_TWIN / 2;
    bool touches = (x < pw) || (y < pw) || (x - pw + WK_TWIN - 1 >= W) || (y - pw + WK_TWIN - 1 >= H);
    int mb = touches ? 0 : (int)mp[pix] - kmin;
    int mg = touches ? 0 : (int)mp[(size_t)W * H + pix] - kmin;
    double pb = sh->pT[0][mb], pg = sh->pT[1][mg];
    int sb = SIGN * kb, sg = SIGN * kg, sr = SIGN * kr;
    Vi[0] += sb; Vi[1] += sg; Vi[2] += sr;
    Vi[3] += sb * kb; Vi[4] += sb * kg; Vi[5] += sb * kr;
    Vi[6] += sg * kg; Vi[7] += sg * kr; Vi[8] += sr * kr;
    double db = (double)sb, dg = (double)sg, dr = (double)sr;
    Vd[0] += (SIGN > 0 ? pb : -pb);
    Vd[1] += (SIGN > 0 ? pg : -pg);
    Vd[2] = fma(db, pb, Vd[2]); Vd[3] = fma(dg, pb, Vd[3]); Vd[4] = fma(dr, pb, Vd[4]);
    Vd[5] = fma(db, pg, Vd[5]); Vd[6] = fma(dg, pg, Vd[6]); Vd[7] = fma(dr, pg, Vd[7]);
  }
  __device__ __forceinline__ void epilogue(int y, int x, int Ncnt, const uint32_t* si, const double* sd) {
    double N = (double)Ncnt, invN = 1.0 / N;
    double M[6], Sd[3], A[6], rdet;
    gf_build_M(si, N, epsN_k * N * N, M, Sd);
    gf_adjugate(M, A, rdet);
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
    double a[3], b;
    gf_solve(A, rdet, Sd, N, invN, sd[0], sd + 2, a, b);
    ab[0 * n_px + pix] = (float)a[0]; ab[1 * n_px + pix] = (float)a[1]; ab[2 * n_px + pix] = (float)a[2]; ab[3 * n_px + pix] = (float)b;
    gf_solve(A, rdet, Sd, N, invN, sd[1], sd + 5, a, b);
    ab[4 * n_px + pix] = (float)a[0]; ab[5 * n_px + pix] = (float)a[1]; ab[6 * n_px + pix] = (float)a[2]; ab[7 * n_px + pix] = (float)b;
  }
  __device__ void finish() {}
};

// block-wide reductions used by the epilogue policies
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

// GF1b: q = (box(a).k + box(b))/N for blue and green -> J (dehazed_BG) + reductions
struct PolGF1b {
  static constexpr int NI = 0, ND = 8;
  struct Shared {
    double nrm[256];
    FrameConst fc;
  };
  GfCommon g; Shared* sh; int W, H, f;
  const uint8_t* img; const float* ab; float* J;
  unsigned long long jmn[2], jmx[2]; long long jsum[2];
  unsigned rmn, rmx; unsigned long long rsum; unsigned nanf;
  __device__ void init(const GfCommon& gc, int frame, Shared* s, int W_, int H_) {
    g = gc; sh = s; W = W_; H = H_; f = frame;
    size_t n_px = (size_t)W * H;
    img = g.src + (size_t)f * n_px * 3;
    ab = g.ab + (size_t)f * 8 * n_px;
    J = g.J + (size_t)f * 2 * n_px;
    if (threadIdx.x == 0) load_frame_const(g.fs[f], sh->fc);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh->nrm[i] = (double)i / (double)sh->fc.range;
    jmn[0] = jmn[1] = ~0ull; jmx[0] = jmx[1] = 0; jsum[0] = jsum[1] = 0;
    rmn = 255; rmx = 0; rsum = 0; nanf = 0;
    __syncthreads();
  }
  template <int SIGN>
  __device__ __forceinline__ void accum(int y, int x, uint32_t*, double* Vd) const {
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      double v = (double)__ldg(ab + k * n_px + pix);
      Vd[k] += (SIGN > 0 ? v : -v);
    }
  }
  __device__ __forceinline__ void epilogue(int y, int x, int Ncnt, const uint32_t*, const double* sd) {
    double invN = 1.0 / (double)Ncnt;
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
    const uint8_t* p = img + pix * 3;
    int kmin = sh->fc.kmin;
    int k[3] = {(int)p[0] - kmin, (int)p[1] - kmin, (int)p[2] - kmin};
    double kd[3] = {(double)k[0], (double)k[1], (double)k[2]};
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double* s = sd + 4 * c;
      double q = (s[0] * kd[0] + s[1] * kd[1] + s[2] * kd[2] + s[3]) * invN;   // guidedfilter.py:100-101
      if (g.dbg_tref && f == 0) g.dbg_tref[(size_t)c * n_px + pix] = q;
      double Bc = sh->fc.B[c];
      double Jv = (sh->nrm[k[c]] - Bc) / q + Bc;                              // BGDehaze.py:53,55
      float Jf = (float)Jv;
      J[(size_t)c * n_px + pix] = Jf;
      double Jr = (double)Jf;
      if (!(fabs(Jr) < 1.0e6)) { nanf |= 1u; Jr = 0.0; }
      unsigned long long key = dkey(Jr);
      jmn[c] = key < jmn[c] ? key : jmn[c];
      jmx[c] = key > jmx[c] ? key : jmx[c];
      jsum[c] += __double2ll_rn(Jr * 4294967296.0);
    }
    rmn = min(rmn, (unsigned)k[2]); rmx = max(rmx, (unsigned)k[2]); rsum += (unsigned)k[2];
  }
  __device__ void finish() {
    FrameState& s = g.fs[f];
    for (int c = 0; c < 2; c++) {
      unsigned long long a = warp_min_u64(jmn[c]), b = warp_max_u64(jmx[c]);
      long long sm = warp_sum_i64(jsum[c]);
      if ((threadIdx.x & 31) == 0) {
        atomicMin(&s.jmin_key[c], a);
        atomicMax(&s.jmax_key[c], b);
        atomicAdd((unsigned long long*)&s.jsum_fix[c], (unsigned long long)sm);
      }
    }
    unsigned a = warp_reduce_min_u32(rmn), b = warp_reduce_max_u32(rmx);
    long long rs = warp_sum_i64((long long)rsum);
    unsigned nf = __reduce_or_sync(0xffffffffu, nanf);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s.rmin, a);
      atomicMax(&s.rmax, b);
      atomicAdd(&s.rsum, (unsigned long long)rs);
      if (nf) atomicOr(&s.nan_flag, nf);
    }
  }
};

// per-pixel guide / S evaluation shared by GF2a, GF2b, E and final
struct ExpShared {
  RedTables rt;
  FrameConst fc;
};
struct PixelExp {
  int gy, gcr, gcb;     // guide = YiCrCb - joint min (k units of normYiCrCb)
  double rest[3];       // restored b, g, r
  int yj;               // Yj - joint min of YjCrCb
};
__device__ __forceinline__ void restored_pixel(const ExpShared* sh, const uint8_t* p, float jb, float jg, double* rest, int* r8, int* i8) {
  int kmin = sh->fc.kmin;
  int kb = (int)p[0] - kmin, kg = (int)p[1] - kmin, kr = (int)p[2] - kmin;
  rest[0] = norm_j(jb, sh->fc, 0);
  rest[1] = norm_j(jg, sh->fc, 1);
  rest[2] = sh->rt.redN[kr];
  r8[0] = trunc_u8(rest[0] * 255.0); r8[1] = trunc_u8(rest[1] * 255.0); r8[2] = sh->rt.red8[kr];
  i8[0] = sh->rt.i8[kb]; i8[1] = sh->rt.i8[kg]; i8[2] = sh->rt.i8[kr];
}
__device__ __forceinline__ void eval_pixel(const ExpShared* sh, const uint8_t* p, float jb, float jg, PixelExp& e) {
  int r8[3], i8[3];
  restored_pixel(sh, p, jb, jg, e.rest, r8, i8);
  int Y, Cr, Cb;
  bgr2ycrcb_u8(i8[0], i8[1], i8[2], Y, Cr, Cb);
  e.gy = Y - sh->fc.yi_min; e.gcr = Cr - sh->fc.yi_min; e.gcb = Cb - sh->fc.yi_min;
  bgr2ycrcb_u8(r8[0], r8[1], r8[2], Y, Cr, Cb);
  e.yj = Y - sh->fc.yj_min;
}
__device__ void exp_shared_init(ExpShared* sh, const FrameState& s, double n_px) {
  if (threadIdx.x == 0) load_frame_const(s, sh->fc);
  __syncthreads();
  build_red_tables(s, sh->fc, n_px, &sh->rt, threadIdx.x, blockDim.x);
  __syncthreads();
}

// GF2a: guide = normYiCrCb (k units), p = S (BGDehaze.py:83)
struct PolGF2a {
  static constexpr int NI = 9, ND = 4;
  typedef ExpShared Shared;
  GfCommon g; Shared* sh; int W, H, f;
  const uint8_t* img; const float* J; float* ab;
  double epsN_k; unsigned nanf;
  __device__ void init(const GfCommon& gc, int frame, Shared* s, int W_, int H_) {
    g = gc; sh = s; W = W_; H = H_; f = frame;
    size_t n_px = (size_t)W * H;
    img = g.src + (size_t)f * n_px * 3;
    J = g.J + (size_t)f * 2 * n_px;
    ab = g.ab + (size_t)f * 8 * n_px;
    exp_shared_init(sh, g.fs[f], (double)n_px);
    double rng = (double)sh->fc.yi_rng;
    epsN_k = g.eps * rng * rng;
    nanf = 0;
  }
  template <int SIGN>
  __device__ __forceinline__ void accum(int y, int x, uint32_t* Vi, double* Vd) {
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
    PixelExp e;
    eval_pixel(sh, img + pix * 3, __ldg(J + pix), __ldg(J + n_px + pix), e);
    double yi = (double)e.gy / (double)sh->fc.yi_rng, yj = (double)e.yj / (double)sh->fc.yj_rng;
    double yi2 = 0.3 * (yi * yi);
    double S = (yj * yi + yi2) / (yj * yj + yi2);
    if (SIGN > 0 && !(S == S)) nanf = 1u;
    int s0 = SIGN * e.gy, s1 = SIGN * e.gcr, s2 = SIGN * e.gcb;
    Vi[0] += s0; Vi[1] += s1; Vi[2] += s2;
    Vi[3] += s0 * e.gy; Vi[4] += s0 * e.gcr; Vi[5] += s0 * e.gcb;
    Vi[6] += s1 * e.gcr; Vi[7] += s1 * e.gcb; Vi[8] += s2 * e.gcb;
    Vd[0] += (SIGN > 0 ? S : -S);
    Vd[1] = fma((double)s0, S, Vd[1]); Vd[2] = fma((double)s1, S, Vd[2]); Vd[3] = fma((double)s2, S, Vd[3]);
  }
  __device__ __forceinline__ void epilogue(int y, int x, int Ncnt, const uint32_t* si, const double* sd) {
    double N = (double)Ncnt, invN = 1.0 / N;
    double M[6], Sd[3], A[6], rdet, a[3], b;
    gf_build_M(si, N, epsN_k * N * N, M, Sd);
    gf_adjugate(M, A, rdet);
    gf_solve(A, rdet, Sd, N, invN, sd[0], sd + 1, a, b);
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
    ab[0 * n_px + pix] = (float)a[0]; ab[1 * n_px + pix] = (float)a[1]; ab[2 * n_px + pix] = (float)a[2]; ab[3 * n_px + pix] = (float)b;
  }
  __device__ void finish() {
    unsigned nf = __reduce_or_sync(0xffffffffu, nanf);
    if ((threadIdx.x & 31) == 0 && nf) atomicOr(&g.fs[f].nan_flag, 1u);
  }
};

// GF2b: refined S -> exposure product -> min/max (BGDehaze.py:84-89)
struct PolGF2b {
  static constexpr int NI = 0, ND = 4;
  typedef ExpShared Shared;
  GfCommon g; Shared* sh; int W, H, f;
  const uint8_t* img; const float* J; const float* ab; float* refS;
  unsigned long long omn, omx; unsigned nanf;
  __device__ void init(const GfCommon& gc, int frame, Shared* s, int W_, int H_) {
    g = gc; sh = s; W = W_; H = H_; f = frame;
    size_t n_px = (size_t)W * H;
    img = g.src + (size_t)f * n_px * 3;
    J = g.J + (size_t)f * 2 * n_px;
    ab = g.ab + (size_t)f * 8 * n_px;
    refS = g.refS + (size_t)f * n_px;
    exp_shared_init(sh, g.fs[f], (double)n_px);
    omn = ~0ull; omx = 0; nanf = 0;
  }
  template <int SIGN>
  __device__ __forceinline__ void accum(int y, int x, uint32_t*, double* Vd) const {
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double v = (double)__ldg(ab + k * n_px + pix);
      Vd[k] += (SIGN > 0 ? v : -v);
    }
  }
  __device__ __forceinline__ void epilogue(int y, int x, int Ncnt, const uint32_t*, const double* sd) {
    double invN = 1.0 / (double)Ncnt;
    size_t n_px = (size_t)W * H, pix = (size_t)y * W + x;
    PixelExp e;
    eval_pixel(sh, img + pix * 3, __ldg(J + pix), __ldg(J + n_px + pix), e);
    double q = (sd[0] * (double)e.gy + sd[1] * (double)e.gcr + sd[2] * (double)e.gcb + sd[3]) * invN;
    float qf = (float)q;
    refS[pix] = qf;
    double qr = (double)qf;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double o = e.rest[c] * qr;
      if (!(o == o)) { nanf = 1u; o = 0.0; }
      unsigned long long key = dkey(o);
      omn = key < omn ? key : omn;
      omx = key > omx ? key : omx;
    }
  }
  __device__ void finish() {
    unsigned long long a = warp_min_u64(omn), b = warp_max_u64(omx);
    unsigned nf = __reduce_or_sync(0xffffffffu, nanf);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&g.fs[f].omin_key, a);
      atomicMax(&g.fs[f].omax_key, b);
      if (nf) atomicOr(&g.fs[f].nan_flag, 1u);
    }
  }
};

// -------------------------------------------------------------------------------------------------
// the march kernel
// -------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void gf_phase1(T* arr, int g16, unsigned hmask) {
  // serial inclusive prefix over this segment's GF_SEG columns, then add the exclusive scan of the
  // 16 segment totals (half-warp shuffle scan) so that arr[] holds the prefix over the whole strip.
  T tot = 0;
#pragma unroll
  for (int i = 0; i < GF_SEG; i++) tot += arr[i * (GF_NSEG + 1) + g16];
  T incl = tot;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    T o = __shfl_up_sync(hmask, incl, d, 16);
    if (g16 >= d) incl += o;
  }
  T run = incl - tot;
#pragma unroll
  for (int i = 0; i < GF_SEG; i++) {
    run += arr[i * (GF_NSEG + 1) + g16];
    arr[i * (GF_NSEG + 1) + g16] = run;
  }
}

template <class P>
__global__ void __launch_bounds__(GF_NT, 2) gf_march_kernel(GfCommon gc, int W, int H, int r, int seg_h) {
  constexpr int NI = P::NI, ND = P::ND, NQ = NI + ND;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_d = reinterpret_cast<double*>(smem_raw);                       // [ND][GF_PITCH]
  uint32_t* s_i = reinterpret_cast<uint32_t*>(s_d + ND * GF_PITCH);        // [NI][GF_PITCH]
  typename P::Shared* sh = reinterpret_cast<typename P::Shared*>(smem_raw + (((size_t)ND * GF_PITCH * 8 + (size_t)NI * GF_PITCH * 4 + 15) & ~(size_t)15));
  P pol;
  pol.init(gc, blockIdx.z, sh, W, H);

  const int t = threadIdx.x;
  const int SW = GF_NT - 2 * r;
  const int xs = blockIdx.x * SW;
  const int x = xs - r + t;
  const bool xin = (x >= 0 && x < W);
  const int ys = blockIdx.y * seg_h, ye = min(ys + seg_h, H);
  const bool is_out = (t >= r) && (t < GF_NT - r) && (x < W);
  const int nx = min(x + r, W - 1) - max(x - r, 0) + 1;
  const int my_slot = (t % GF_SEG) * (GF_NSEG + 1) + t / GF_SEG;
  const int hi_c = t + r, lo_c = t - r - 1;
  const int hi_slot = (hi_c % GF_SEG) * (GF_NSEG + 1) + hi_c / GF_SEG;
  const int lo_slot = lo_c >= 0 ? (lo_c % GF_SEG) * (GF_NSEG + 1) + lo_c / GF_SEG : 0;

  uint32_t Vi[NI > 0 ? NI : 1];
  double Vd[ND > 0 ? ND : 1];
#pragma unroll
  for (int k = 0; k < NI; k++) Vi[k] = 0;
#pragma unroll
  for (int k = 0; k < ND; k++) Vd[k] = 0.0;

  for (int yin = ys - r; yin < ye + r; ++yin) {
    if (xin) {
      if (yin >= 0 && yin < H) pol.template accum<+1>(yin, x, Vi, Vd);
      int yl = yin - 2 * r - 1;
      if (yl >= ys - r && yl >= 0) pol.template accum<-1>(yl, x, Vi, Vd);
    }
    int yo = yin - r;
    if (yo < ys) continue;  // warm-up rows (uniform across the CTA)
#pragma unroll
    for (int k = 0; k < NI; k++) s_i[k * GF_PITCH + my_slot] = Vi[k];
#pragma unroll
    for (int k = 0; k < ND; k++) s_d[k * GF_PITCH + my_slot] = Vd[k];
    __syncthreads();
    if (t < NQ * 16) {
      int q = t >> 4, g16 = t & 15;
      unsigned hmask = 0xffffu << (t & 16);  // the two half-warps scan different quantities
      if (q < NI) gf_phase1<uint32_t>(s_i + q * GF_PITCH, g16, hmask);
      else gf_phase1<double>(s_d + (q - NI) * GF_PITCH, g16, hmask);
    }
    __syncthreads();
    if (is_out) {
      uint32_t si[NI > 0 ? NI : 1];
      double sd[ND > 0 ? ND : 1];
#pragma unroll
      for (int k = 0; k < NI; k++) si[k] = s_i[k * GF_PITCH + hi_slot] - (lo_c >= 0 ? s_i[k * GF_PITCH + lo_slot] : 0u);
#pragma unroll
      for (int k = 0; k < ND; k++) sd[k] = s_d[k * GF_PITCH + hi_slot] - (lo_c >= 0 ? s_d[k * GF_PITCH + lo_slot] : 0.0);
      int ny = min(yo + r, H - 1) - max(yo - r, 0) + 1;
      pol.epilogue(yo, x, ny * nx, si, sd);
    }
    __syncthreads();
  }
  pol.finish();
}

template <class P>
static size_t gf_smem_bytes() {
  return (((size_t)P::ND * GF_PITCH * 8 + (size_t)P::NI * GF_PITCH * 4 + 15) & ~(size_t)15) + sizeof(typename P::Shared);
}

template <class P>
static int gf_launch(uwip_ctx* ctx, const char* tag, const GfCommon& gc, int n, int W, int H, int r) {
  int SW = GF_NT - 2 * r;
  int strips = cdiv(W, SW);
  // split rows into segments until there are ~2 waves of CTAs (2 CTAs per SM resident)
  int target = ctx->sm_count * 4;
  int segs = cdiv(target, strips * n);
  int max_segs = H / (4 * r + 2) > 1 ? H / (4 * r + 2) : 1;  // keep the 2r warm-up rows below ~1/3 of the work
  if (segs > max_segs) segs = max_segs;
  if (segs < 1) segs = 1;
  int seg_h = cdiv(H, segs);
  segs = cdiv(H, seg_h);
  size_t smem = gf_smem_bytes<P>();
  static bool attr_done = false;  // per template instantiation
  if (!attr_done) {
    UWIP_CUDA(ctx, cudaFuncSetAttribute(gf_march_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid(strips, segs, n);
  UWIP_LAUNCH(ctx, tag, gf_march_kernel<P>, grid, GF_NT, smem, gc, W, H, r, seg_h);
  return UWIP_OK;
}

// -------------------------------------------------------------------------------------------------
// E: restored -> R8, I8 -> YCrCb joint min / max (BGDehaze.py:75-80)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exposure_minmax_kernel(GfCommon g, int W, int H, double* dbg_restored) {
  __shared__ ExpShared sh;
  int f = blockIdx.y;
  size_t n_px = (size_t)W * H;
  exp_shared_init(&sh, g.fs[f], (double)n_px);
  const uint8_t* img = g.src + (size_t)f * n_px * 3;
  const float* J = g.J + (size_t)f * 2 * n_px;
  unsigned imn = 255, imx = 0, jmn = 255, jmx = 0;
  for (size_t pix = (size_t)blockIdx.x * 256 + threadIdx.x; pix < n_px; pix += (size_t)gridDim.x * 256) {
    double rest[3];
    int r8[3], i8[3];
    restored_pixel(&sh, img + pix * 3, __ldg(J + pix), __ldg(J + n_px + pix), rest, r8, i8);
    if (dbg_restored && f == 0) { dbg_restored[pix * 3] = rest[0]; dbg_restored[pix * 3 + 1] = rest[1]; dbg_restored[pix * 3 + 2] = rest[2]; }
    int Y, Cr, Cb;
    bgr2ycrcb_u8(i8[0], i8[1], i8[2], Y, Cr, Cb);
    imn = min(imn, (unsigned)imin3(Y, Cr, Cb)); imx = max(imx, (unsigned)imax3(Y, Cr, Cb));
    bgr2ycrcb_u8(r8[0], r8[1], r8[2], Y, Cr, Cb);
    jmn = min(jmn, (unsigned)imin3(Y, Cr, Cb)); jmx = max(jmx, (unsigned)imax3(Y, Cr, Cb));
  }
  imn = warp_reduce_min_u32(imn); imx = warp_reduce_max_u32(imx);
  jmn = warp_reduce_min_u32(jmn); jmx = warp_reduce_max_u32(jmx);
  if ((threadIdx.x & 31) == 0) {
    FrameState& s = g.fs[f];
    atomicMin(&s.yi_min, imn); atomicMax(&s.yi_max, imx);
    atomicMin(&s.yj_min, jmn); atomicMax(&s.yj_max, jmx);
  }
}

// final: (OutputExp - min)/(max - min) * 255 -> rint -> saturate (BGDehaze.py:88-89, main.py:19)
__global__ void __launch_bounds__(256) final_kernel(GfCommon g, int W, int H, uint8_t* __restrict__ dst, double* dbg_out, int32_t* __restrict__ flags) {
  __shared__ ExpShared sh;
  int f = blockIdx.y;
  size_t n_px = (size_t)W * H;
  exp_shared_init(&sh, g.fs[f], (double)n_px);
  const FrameState& s = g.fs[f];
  double omn = dunkey(s.omin_key), den = dunkey(s.omax_key) - omn;
  bool nan_frame = s.nan_flag != 0;
  if (flags && blockIdx.x == 0 && threadIdx.x == 0) flags[f] = nan_frame ? UWIP_FRAME_NAN : 0;
  const uint8_t* img = g.src + (size_t)f * n_px * 3;
  const float* J = g.J + (size_t)f * 2 * n_px;
  const float* refS = g.refS + (size_t)f * n_px;
  uint8_t* out = dst + (size_t)f * n_px * 3;
  for (size_t pix = (size_t)blockIdx.x * 256 + threadIdx.x; pix < n_px; pix += (size_t)gridDim.x * 256) {
    double rest[3];
    int r8[3], i8[3];
    restored_pixel(&sh, img + pix * 3, __ldg(J + pix), __ldg(J + n_px + pix), rest, r8, i8);
    double q = (double)__ldg(refS + pix);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double o = (rest[c] * q - omn) / den;
      if (nan_frame) o = __longlong_as_double(0x7ff8000000000000ll);
      if (dbg_out && f == 0) dbg_out[pix * 3 + c] = o;
      double v = o * 255.0;
      int b = 0;
      if (v == v && fabs(v) < 2.0e9) b = min(max(__double2int_rn(v), 0), 255);
      out[pix * 3 + c] = (uint8_t)b;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// driver
// -------------------------------------------------------------------------------------------------
int dehaze_frames_dev(uwip_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, int n, int W, int H, const uwip_dehaze_params& p,
                      bool minmax_done, FrameState* fs, DehazeDebug* dbg, int32_t* d_flags) {
  UWIP_REQUIRE(ctx, n >= 1 && W >= 1 && H >= 1, "bad size");
  UWIP_REQUIRE(ctx, p.window >= 1 && p.window <= WK_MAXWIN, "window must be 1..33");
  UWIP_REQUIRE(ctx, p.radius >= 1 && 2 * p.radius <= GF_NT - 64, "radius must be 1..160");
  UWIP_REQUIRE(ctx, (size_t)W * H < (1ull << 31), "frame too large");
  UWIP_REQUIRE(ctx, !dbg || n == 1, "stage outputs are single-frame");
  size_t n_px = (size_t)W * H;
  if (!minmax_done) {
    int want = (int)std::min<size_t>((n_px * 3 / 16 + 255) / 256, 1u << 16);
    int gxm = std::max(1, std::min(want, ctx->sm_count * 8 / n + 1));
    dim3 grid(gxm, n);
    UWIP_LAUNCH(ctx, "dz_minmax", minmax_kernel, grid, 256, 0, d_src, n_px * 3, fs);
  }
  uint8_t* d_m = (uint8_t*)uwip_slot(ctx, SLOT_MPLANES, (size_t)n * 2 * n_px);
  dim3 gridw(cdiv(W, WK_TX), cdiv(H, WK_TY), n);
  int n_part = gridw.x * gridw.y;
  ArgPartial* d_part = (ArgPartial*)uwip_slot(ctx, SLOT_PARTIALS, (size_t)n * n_part * sizeof(ArgPartial));
  float* d_ab = (float*)uwip_slot(ctx, SLOT_AB, (size_t)n * 8 * n_px * sizeof(float));
  float* d_J = (float*)uwip_slot(ctx, SLOT_J, (size_t)n * 2 * n_px * sizeof(float));
  float* d_refS = (float*)uwip_slot(ctx, SLOT_REFS, (size_t)n * n_px * sizeof(float));
  if (!d_m || !d_part || !d_ab || !d_J || !d_refS) return UWIP_ERR_NOMEM;
  {
    const int wmax = p.window, wmin = WK_TWIN;
    int PL = std::max(wmax / 2, wmin / 2), PR = std::max(wmax - 1 - wmax / 2, wmin - 1 - wmin / 2);
    int RW = WK_TX + PL + PR, RH = WK_TY + PL + PR;
    size_t smem = ((size_t)2 * RW * RH + (size_t)3 * RH * WK_TX) * 4;
    static size_t attr = 0;
    if (smem > attr) {
      UWIP_CUDA(ctx, cudaFuncSetAttribute(window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    UWIP_LAUNCH(ctx, "dz_window", window_kernel, gridw, WK_THREADS, smem, d_src, W, H, wmax, fs, d_m, d_part);
    UWIP_LAUNCH(ctx, "dz_bglight", bglight_finish_kernel, n, 256, 0, d_src, W, H, d_part, n_part, fs, 1, wmax == WK_TWIN ? 1 : 0);
    if (wmax != WK_TWIN) {
      // refined_t() is always called without w (BGDehaze.py:52): the transmission uses the
      // background light of a 15x15 window even when Background_light for J uses another w
      UWIP_LAUNCH(ctx, "dz_window", window_kernel, gridw, WK_THREADS, smem, d_src, W, H, WK_TWIN, fs, d_m, d_part);
      UWIP_LAUNCH(ctx, "dz_bglight", bglight_finish_kernel, n, 256, 0, d_src, W, H, d_part, n_part, fs, 0, 1);
    }
  }
  if (dbg && dbg->stop_after == 1) return UWIP_OK;
  if (dbg && dbg->t_raw) {
    dim3 g2(cdiv(W, 256), H);
    UWIP_LAUNCH(ctx, "dz_traw", traw_kernel, g2, 256, 0, d_m, W, H, fs, dbg->t_raw);
  }
  if (dbg && dbg->stop_after == 2) return UWIP_OK;

  GfCommon gc;
  gc.src = d_src; gc.mplanes = d_m; gc.ab = d_ab; gc.J = d_J; gc.refS = d_refS; gc.fs = fs;
  gc.eps = p.eps; gc.tmin = p.tmin; gc.dbg_tref = dbg ? dbg->t_ref : nullptr;
  UWIP_CHECK(gf_launch<PolGF1a>(ctx, "dz_gf1a", gc, n, W, H, p.radius));
  UWIP_CHECK(gf_launch<PolGF1b>(ctx, "dz_gf1b", gc, n, W, H, p.radius));
  if (dbg && dbg->stop_after == 3) return UWIP_OK;
  int gx = std::max(1, std::min((int)((n_px + 2047) / 2048), std::max(1, ctx->sm_count * 8 / n)));
  dim3 grid_e(gx, n);
  UWIP_LAUNCH(ctx, "dz_exposure_minmax", exposure_minmax_kernel, grid_e, 256, 0, gc, W, H, dbg ? dbg->restored : (double*)nullptr);
  if (dbg && dbg->stop_after == 4) return UWIP_OK;
  UWIP_CHECK(gf_launch<PolGF2a>(ctx, "dz_gf2a", gc, n, W, H, p.radius));
  UWIP_CHECK(gf_launch<PolGF2b>(ctx, "dz_gf2b", gc, n, W, H, p.radius));
  UWIP_LAUNCH(ctx, "dz_final", final_kernel, grid_e, 256, 0, gc, W, H, d_dst, dbg ? dbg->out : (double*)nullptr, d_flags);
  return UWIP_OK;
}

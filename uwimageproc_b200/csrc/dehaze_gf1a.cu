// dehaze_gf1a.cu - the GF1a march (17 box moments of the transmission filter + per-pixel 3x3 solve, guidedfilter.py:23-75)
// compiled with the narrow strip layout of gfpipe.cuh: 128 quads per strip, four ACC warps in one warpgroup with 176
// registers, the four auxiliary warps in their own warpgroup with 88, SOLVE with 120.  In the wide layout (160 quads: ACC 152 /
// SOLVE 104, the auxiliary warps sharing the ACC warpgroup's count) both roles of this kernel spilled loop invariants to
// local memory and the reloads sat on the critical path of every row; the other marches have no such pressure and lose
// 7 - 12 % to the 9/8 more strips of the narrow layout (profiles/r2_summary.md), so they stay in dehaze.cu.
#define GP_NARROW 1
#include "gfmarch.cuh"

int dehaze_gf1a_launch(uwip_ctx* ctx, const GfCommon& gc, int n, int W, int H, int r) {
  return gp_launch<PipGF1a>(ctx, "dz_gf1a", FUNC_GF1A, gc, n, W, H, r);
}

int dehaze_gf1a_strips(int w) {
  GfGeom g = gf_geometry(w, 4 * 40 + 2, 40);
  return cdiv(w, g.SW);
}

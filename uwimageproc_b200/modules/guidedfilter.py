"""Mirror of modules/bgdehaze/guidedfilter.py (reference function names kept): boxfilter and guided_filter as stage
entry points.  Inside the chain both are fused into the guided-filter marches (csrc/gfpipe.cuh); these wrappers exist so
that every stage of SURVEY 8a (D4, D5) can be checked on its own.  The arithmetic runs in libuwip.so; no numpy fallback.
"""
import numpy as np

from ..api import default_context


def boxfilter(I, r):  # guidedfilter.py:23
    """(2r+1)^2 window sum of a float64 plane, windows truncated at the borders."""
    return default_context().boxfilter(I, r)


def _guide_u8(I):
    """(guide8, range) with I == guide8 / range exactly: every guide of the path is an 8-bit image divided by its joint
    range (normI: bgdehaze/main.py:17, normYiCrCb: BGDehaze.py:79-80)."""
    a = np.asarray(I)
    if a.dtype == np.uint8:
        lo, hi = int(a.min()), int(a.max())
        return np.ascontiguousarray(a - lo), max(hi - lo, 1)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an H x W x 3 guide")
    a = a.astype(np.float64, copy=False)
    if not (np.nanmin(a) >= 0.0 and np.nanmax(a) <= 1.0):
        raise ValueError("the guide must be normalised to [0, 1]")
    for rng in range(255, 0, -1):
        k = a * rng
        kr = np.rint(k)
        if np.abs(k - kr).max() < 1e-9 * rng:
            return kr.astype(np.uint8), rng
    raise ValueError("the guide does not come from an 8-bit image; the CUDA path is defined for 8-bit guides")


def guided_filter(I, p, r=40, eps=1e-3):  # guidedfilter.py:54
    """Colour-guide guided filter (He et al.): I = H x W x 3 guide in [0, 1], p = H x W signal in [0, 1.6]."""
    g8, rng = _guide_u8(I)
    return default_context().guided_filter_u8(g8, rng, p, r, eps)

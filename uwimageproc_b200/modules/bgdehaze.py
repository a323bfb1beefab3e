"""Mirror of modules/bgdehaze/BGDehaze.py + main.py (reference function names kept).

Every function takes what the reference takes - `normI`, the float64 H x W x 3 BGR image in [0, 1]
produced by `(I - I.min()) / (I.max() - I.min())` from an 8-bit frame (bgdehaze/main.py:17) - or the
8-bit frame itself, and returns what the reference returns.  The arithmetic runs in libuwip.so
(uwip_*_bgr8 entry points of include/uwip.h); there is no numpy fallback.

guided_filter / boxfilter (guidedfilter.py:23,54) are internal helpers of refined_t and
adaptiveExp_map in the reference; here they are fused into those kernels and not exported.
"""
import numpy as np

from ..api import default_context


def _as_u8(img):
    """8-bit frame whose main.py:17 normalisation reproduces `img` exactly."""
    a = np.asarray(img)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected H x W x 3 (BGR)")
    a = a.astype(np.float64, copy=False)
    if not (np.nanmin(a) == 0.0 and np.nanmax(a) == 1.0):
        raise ValueError("normI must be the min-max normalised frame of bgdehaze/main.py:17 (min 0, max 1)")
    for rng in range(255, 0, -1):  # the 8-bit range max-min that generated normI
        k = a * rng
        kr = np.rint(k)
        if np.abs(k - kr).max() < 1e-9 * rng:
            return kr.astype(np.uint8)
    raise ValueError("normI does not come from an 8-bit frame; the CUDA path is defined for 8-bit sources")


def Background_light(normI, w=15):  # BGDehaze.py:14
    return default_context().background_light(_as_u8(normI), w)[0]


def transmission_map(normI, w=15):  # BGDehaze.py:28
    tb, tg = default_context().transmission(_as_u8(normI), w)
    return np.stack([tb, tg], axis=2)


def refined_t(normI, w=15):  # BGDehaze.py:39 (w is accepted and ignored exactly like the reference's caller, :52)
    ctx = default_context()
    tb, tg = ctx.refined_transmission(_as_u8(normI), ctx.dehaze_params(window=15))
    return tb, tg


def RC_correction(normI, w=15):  # BGDehaze.py:59
    ctx = default_context()
    return ctx.rc_correction(_as_u8(normI), ctx.dehaze_params(window=w))


def dehazed_BG(normI, w=15):  # BGDehaze.py:50 -> (normJb, normJg)
    r = RC_correction(normI, w)
    return r[..., 0].copy(), r[..., 1].copy()


def adaptiveExp_map(normI, w=15):  # BGDehaze.py:71
    ctx = default_context()
    return ctx.bgdehaze(_as_u8(normI), ctx.dehaze_params(window=w), return_float=True)[1]


def generate_results(I, w=15):
    """bgdehaze/main.py:14-20 without the file I/O: 8-bit BGR frame in, the bytes imwrite would encode out."""
    ctx = default_context()
    return ctx.bgdehaze(_as_u8(I), ctx.dehaze_params(window=w))

"""uwimageproc_b200: B200-native histretch -> aclahe -> bgdehaze behind the C ABI of include/uwip.h.

Python here is host-side plumbing that mirrors the reference's own Python/C++ entry points
(modules/common/preprocessing, modules/aclahe/python, modules/bgdehaze); the arithmetic runs in the
hand-written sm_100a kernels of libuwip.so.  There is no CPU fallback.
"""
from ._lib import UwipError, load  # noqa: F401
from .api import Context, default_context, host_checksum  # noqa: F401

__all__ = ["Context", "default_context", "UwipError", "load", "host_checksum"]

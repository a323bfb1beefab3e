"""Mirror of modules/common/preprocessing.{h,cpp} for Python callers (names as in the C++ header)."""
import numpy as np

from .. import _lib
from ..api import default_context


def getHistogram(img):  # preprocessing.h:38 -> float32[256] (the CV_32F column calcHist returns)
    return default_context().histogram(img)


def imgChannelStretch(imgOriginal, imgStretched=None, lowerPercentile=0, higherPercentile=100):  # preprocessing.h:66
    """In place when imgStretched is the same array, as every reference caller does (histretch.cpp:236,247)."""
    out = default_context().channel_stretch(imgOriginal, lowerPercentile, higherPercentile)
    if imgStretched is not None:
        np.copyto(imgStretched, out)
        return imgStretched
    return out


imgChannelStretchGPU = imgChannelStretch  # preprocessing.h:96: same contract on the GPU


def numChannel(c):  # preprocessing.h:112
    return int(_lib.load().uwip_num_channel(c.encode()[:1]))


def numSpace(c):  # preprocessing.h:115
    return int(_lib.load().uwip_num_space(c.encode()[:1]))


def histretch(src, channels, lo=2, hi=98, literal=False):
    """The -c=<letters> channel loop of histretch.cpp:219-254 on an 8-bit BGR frame (2/98: histretch.cpp:154)."""
    return default_context().histretch(src, channels, lo, hi, "literal" if literal else "intended")

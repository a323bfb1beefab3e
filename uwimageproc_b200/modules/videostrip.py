"""modules/videostrip: the one function that composes with the enhancement chain.
Same name and argument as the reference; the work runs in libuwip.so on the GPU (no CPU fallback)."""
from ..api import default_context


def calcBlur(frame):  # videostrip.hpp:98, videostrip.cpp:170-184
    """Standard deviation of the 8-bit aperture-3 Laplacian of the grey frame (float32): low = blurred."""
    return default_context().calc_blur(frame)


def calcBlurGPU(frame):  # videostrip.hpp:106, videostrip.cpp:39-60
    """The cv::cuda twin asks for the aperture-1 kernel [0 1 0; 1 -4 1; 0 1 0] (videostrip.cpp:48), so it does not
    return calcBlur's number; cv::cuda cannot run in this image, the kernel is pinned against cv2.Laplacian(ksize=1)."""
    return default_context().calc_blur(frame, aperture=1)

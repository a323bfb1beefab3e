"""Host-side mirrors of the reference's Python modules for the hot path (same function names, argument
meaning and array conventions), calling libuwip.so through the C ABI.  No numerics happen here apart
from the 49-point curve fit of ParametrosACLAHE, which the reference also does on the host."""

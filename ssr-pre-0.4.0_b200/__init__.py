"""B200-native drop-in for the SoundScape Renderer's apf::conv partitioned-convolution
path.  The product is lib/libssr_b200.so (CUDA kernels for sm_100a + C ABI,
include/ssr_b200.h); this Python package is only the ctypes binding used by the
tests, bench.py and the multi-GPU driver.  There is no CPU fallback: importing
`capi` fails loudly when the shared library has not been built.

The directory name is not a Python identifier; load it with
`__graft_entry__.load_package()` which registers it as `ssr_b200`.
"""

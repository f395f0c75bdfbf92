"""jeicyboodsp_b200 -- B200-native (sm_100a) frame-wise spectral hot path of phoenix163/JeicybooDSP.

The product is `libjdsp.so` (hand-written CUDA behind the C ABI in include/jdsp.h); this package only
holds its sources (csrc/), the build recipe, a ctypes binding used by tests and bench.py, and the
synthetic-workload generators.  Nothing here computes on the CPU.
"""
from .binding import SS, WIENER, Context, JdspError, Library  # noqa: F401

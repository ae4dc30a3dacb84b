"""hemocell_b200: B200-native implementation of HemoCell's per-timestep IB-LBM hot path.

The product is libhemocell_gpu.so (hand-written sm_100a kernels behind the C ABI of
include/hemocell_gpu.h) plus the C++ host-side mirror of the reference interface under
hemocell_b200/host/.  `lib` is the ctypes binding the test/bench harness uses.
There is no CPU fallback: importing `lib.load()` without the built library raises.
"""
from . import lib  # noqa: F401

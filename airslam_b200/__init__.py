"""airslam_b200 -- B200-native learned front-end for AirSLAM (detect + match), behind a C ABI.

The product is `libairfe.so` (hand-written sm_100a CUDA, see csrc/) and the C++ class surfaces in include/.
This Python package is only the ctypes binding used by tests and bench.py; it contains no compute.
"""
from .capi import lib, check, AirfeError  # noqa: F401

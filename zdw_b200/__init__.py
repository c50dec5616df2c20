"""zdw_b200 -- B200 (sm_100a) implementation of the adobe/zdw hot path.

The product is native: `libzdw_b200.so` (hand-written CUDA kernels behind the C ABI declared in
include/zdw_b200.h) plus the C++ host classes/CLIs under zdw_b200/host.  This Python package only
holds the ctypes binding the test-suite and bench.py use to call that ABI; it contains no compute
and no CPU fallback.
"""
from .capi import (  # noqa: F401
    Context,
    EncodedBlock,
    DecodedBlock,
    ZdwError,
    lib_path,
    load_library,
)

__all__ = ["Context", "EncodedBlock", "DecodedBlock", "ZdwError", "lib_path", "load_library"]

"""rajaperf_b200 -- the Base_B200 kernel variant for the RAJA Performance Suite hot path.

Hand-written sm_100a CUDA kernels behind a C ABI (include/rpb200.h, built into
rajaperf_b200/lib/librpb200.so), a ctypes binding of that ABI (cabi.py), and the C++ suite
harness that mirrors the reference's KernelBase API (suite/).  No CPU / PyTorch fallback.
"""
from . import cabi  # noqa: F401
from .cabi import Context, RPB200Error  # noqa: F401

__all__ = ["cabi", "Context", "RPB200Error"]

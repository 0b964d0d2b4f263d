"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None
_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_L = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ORACLE_DIR, "rpb_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    L = ctypes.CDLL(LIB)
    sig = {
        "orc_reset_init_count": (None, []),
        "orc_init_real": (None, [_D, c_int64]),
        "orc_init_const": (None, [_D, c_int64, c_double]),
        "orc_init_rand_value": (None, [_D, c_int64]),
        "orc_init_rand_sign": (None, [_D, c_int64]),
        "orc_init_int": (None, [_I, c_int64]),
        "orc_init_scalar": (c_double, []),
        "orc_checksum_out": (None, [_D, c_int64, c_double, c_void_p]),
        "orc_stream_copy": (None, [_D, _D, c_int64]),
        "orc_stream_mul": (None, [_D, _D, c_double, c_int64]),
        "orc_stream_add": (None, [_D, _D, _D, c_int64]),
        "orc_stream_triad": (None, [_D, _D, _D, c_double, c_int64]),
        "orc_stream_dot": (c_double, [_D, _D, c_int64, c_double]),
        "orc_reduce_sum": (c_double, [_D, c_int64, c_double]),
        "orc_scan_exclusive": (None, [_D, _D, c_int64]),
        "orc_sort": (None, [_D, c_int64]),
        "orc_sort_pairs": (None, [_D, _D, c_int64]),
        "orc_mass3dpa": (None, [_D, _D, _D, _D, _D, c_int64]),
        "orc_diffusion3dpa": (None, [_D, _D, _D, _D, _D, c_int64, c_int]),
        "orc_diffusion3dpa_tables": (None, [_D, _D, _D, _D]),
        "orc_convection3dpa": (None, [_D, _D, _D, _D, _D, _D, c_int64]),
        "orc_ltimes": (None, [_D, _D, _D, c_int64, c_int64, c_int64, c_int64]),
        "orc_halo_grid_dims": (None, [c_int64, _L]),
        "orc_halo_extent_len": (c_int64, [c_int, c_int, c_int64, _L]),
        "orc_halo_make_list": (None, [c_int, c_int, c_int64, _L, _I]),
        "orc_halo_neighbors": (None, [c_int, _I, _I, _I, _I]),
        "orc_halo_pack": (None, [_D, _I, _D, c_int64]),
        "orc_halo_unpack": (None, [_D, _I, _D, c_int64]),
        "orc_checksum_int_out": (None, [_I, c_int64, c_double, c_void_p]),
        "orc_indexlist": (c_int64, [_D, _I, c_int64]),
        "orc_indexlist_3loop": (c_int64, [_D, _I, c_int64]),
        "orc_polybench_gemm": (None, [_D, _D, _D, c_int64, c_int64, c_int64, c_double, c_double]),
        "orc_polybench_gemm_dims": (None, [c_int64, _L, _L, _L]),
        "orc_kat": (c_int, [c_char_p, c_int64, c_int, c_void_p, c_void_p]),
        "orc_omp_threads": (c_int, []),
        "orc_stream_copy_omp": (None, [_D, _D, c_int64]),
        "orc_stream_mul_omp": (None, [_D, _D, c_double, c_int64]),
        "orc_stream_add_omp": (None, [_D, _D, _D, c_int64]),
        "orc_stream_triad_omp": (None, [_D, _D, _D, c_double, c_int64]),
        "orc_stream_dot_omp": (c_double, [_D, _D, c_int64, c_double]),
        "orc_reduce_sum_omp": (c_double, [_D, c_int64, c_double]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def checksum(arr: np.ndarray, scale: float = 1.0) -> np.longdouble:
    """Reference checksum (DataUtils.cpp:600-621) as an 80-bit long double."""
    a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
    out = np.zeros(1, dtype=np.longdouble)
    lib().orc_checksum_out(a, a.size, scale, out.ctypes.data_as(c_void_p))
    return out[0]


def checksum_int(arr: np.ndarray, scale: float = 1.0) -> np.longdouble:
    """The Int_ptr overload (DataUtils.cpp:623-629)."""
    a = np.ascontiguousarray(arr, dtype=np.int32).reshape(-1)
    out = np.zeros(1, dtype=np.longdouble)
    lib().orc_checksum_int_out(a, a.size, scale, out.ctypes.data_as(c_void_p))
    return out[0]


def kat(kernel: str, size: int = 0, reps: int = 1, iparams=None) -> np.longdouble:
    """Whole-kernel Base_Seq checksum: setUp -> reps -> updateChecksum."""
    out = np.zeros(1, dtype=np.longdouble)
    ip = None
    if iparams is not None:
        ip = np.ascontiguousarray(iparams, dtype=np.int32)
    rc = lib().orc_kat(kernel.encode(), size, reps,
                       ip.ctypes.data_as(c_void_p) if ip is not None else None,
                       out.ctypes.data_as(c_void_p))
    if rc != 0:
        raise KeyError(kernel)
    return out[0]


def init_real(n: int, count: int = 0) -> np.ndarray:
    """initData(Real_ptr) as the `count`-th init call after a reset (factor 0.2 if even)."""
    L = lib()
    L.orc_reset_init_count()
    dummy = np.zeros(1)
    for _ in range(count):
        L.orc_init_const(dummy, 0, 0.0)
    a = np.empty(n, dtype=np.float64)
    L.orc_init_real(a, n)
    return a


def init_rand_value(n: int) -> np.ndarray:
    a = np.empty(n, dtype=np.float64)
    lib().orc_init_rand_value(a, n)
    return a

"""ctypes binding of the Base_B200 C ABI (include/rpb200.h).

This is the Python-side stub a maintainer would write against ``librpb200.so``; it adds
nothing of its own.  Device memory comes from the caller (torch tensors in the tests and
in bench.py: ``tensor.data_ptr()``), streams from ``torch.cuda.current_stream().cuda_stream``.

There is NO fallback: if the shared library is missing, or no sm_100 device is usable,
loading / ``Context()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t,
                    c_ubyte, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librpb200.so")


class RPB200Error(RuntimeError):
    pass


class HaloSeg(Structure):
    """struct rpb200_halo_seg (include/rpb200.h)."""
    _fields_ = [("buffer", c_void_p), ("list", c_void_p), ("var", c_void_p),
                ("len", c_int64), ("msg", c_int), ("flags", c_int)]


# name -> (restype, argtypes); must list every function declared in include/rpb200.h
_P = c_void_p
SIGNATURES = {
    "rpb200_create": (c_int, [c_int, POINTER(_P)]),
    "rpb200_destroy": (None, [_P]),
    "rpb200_error_string": (c_char_p, [c_int]),
    "rpb200_sm_count": (c_int, [_P]),
    "rpb200_version": (c_char_p, []),
    "rpb200_stream_attach": (c_int, [_P, _P]),
    "rpb200_stream_detach": (c_int, [_P, _P]),
    "rpb200_set_tuning": (c_int, [_P, c_char_p, c_int, c_int, c_int]),
    "rpb200_reset_tuning": (c_int, [_P, c_char_p]),
    "rpb200_get_tuning": (c_int, [_P, c_char_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rpb200_stream_copy": (c_int, [_P, _P, _P, c_int64, _P]),
    "rpb200_stream_mul": (c_int, [_P, _P, _P, c_double, c_int64, _P]),
    "rpb200_stream_add": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "rpb200_stream_triad": (c_int, [_P, _P, _P, _P, c_double, c_int64, _P]),
    "rpb200_stream_dot": (c_int, [_P, _P, _P, c_int64, c_double, _P, c_int, _P]),
    "rpb200_reduce_sum": (c_int, [_P, _P, c_int64, c_double, _P, _P]),
    "rpb200_scan_exclusive": (c_int, [_P, _P, _P, c_int64, _P]),
    "rpb200_scan_reserve": (c_int, [_P, c_int64]),
    "rpb200_sort_scratch_bytes": (c_size_t, [c_int64, c_int]),
    "rpb200_sort_keys_f64": (c_int, [_P, _P, c_int64, _P, c_size_t, _P]),
    "rpb200_sort_pairs_f64": (c_int, [_P, _P, _P, c_int64, _P, c_size_t, _P]),
    "rpb200_mass3dpa": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, _P]),
    "rpb200_diffusion3dpa": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "rpb200_convection3dpa": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "rpb200_ltimes": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, _P]),
    "rpb200_indexlist": (c_int, [_P, _P, _P, c_int64, _P, _P]),
    "rpb200_indexlist_reserve": (c_int, [_P, c_int64]),
    "rpb200_memset_f64": (c_int, [_P, _P, c_double, c_int64, _P]),
    "rpb200_polybench_gemm": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_double, c_double, _P]),
    "rpb200_halo_chunk": (c_int, []),
    "rpb200_halo_worklist_create": (c_int, [_P, _P, c_int, POINTER(_P)]),
    "rpb200_halo_worklist_update": (c_int, [_P, _P, c_int, _P]),
    "rpb200_halo_worklist_destroy": (None, [_P]),
    "rpb200_halo_pack": (c_int, [_P, _P, _P]),
    "rpb200_halo_unpack": (c_int, [_P, _P, _P]),
    "rpb200_halo_pack_unpack": (c_int, [_P, _P, _P, _P]),
    "rpb200_debug_halo_units": (c_int, [POINTER(c_int64), POINTER(c_int), POINTER(c_int), POINTER(c_int), c_int,
                                        POINTER(c_int64), POINTER(c_int), POINTER(c_int), POINTER(c_int), c_int, c_int,
                                        POINTER(c_int), c_int, POINTER(c_int), c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rpb200_halo_grid_dims": (None, [c_int64, POINTER(c_int64)]),
    "rpb200_halo_plan_create": (c_int, [_P, POINTER(c_int64), c_int64, c_int, c_int, POINTER(c_int), POINTER(_P)]),
    "rpb200_halo_plan_destroy": (None, [_P]),
    "rpb200_halo_plan_var_size": (c_int64, [_P]),
    "rpb200_halo_plan_neighbor": (c_int, [_P, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                          POINTER(c_int64), POINTER(c_int64), POINTER(_P), POINTER(_P)]),
    "rpb200_halo_plan_bind": (c_int, [_P, POINTER(_P), POINTER(_P), POINTER(_P)]),
    "rpb200_halo_plan_pack": (c_int, [_P, _P]),
    "rpb200_halo_plan_unpack": (c_int, [_P, _P]),
    "rpb200_halo_plan_pack_unpack": (c_int, [_P, _P]),
    "rpb200_halo_exchange_window": (c_int, [_P, POINTER(_P), POINTER(_P), POINTER(c_size_t), POINTER(c_ubyte)]),
    "rpb200_halo_exchange_connect": (c_int, [_P, c_int, POINTER(c_ubyte)]),
    "rpb200_halo_exchange_connect_ptrs": (c_int, [_P, c_int, POINTER(_P)]),
    "rpb200_halo_exchange_pack": (c_int, [_P, _P]),
    "rpb200_halo_exchange_unpack": (c_int, [_P, _P]),
    "rpb200_halo_exchange_pack_seg": (c_int, [_P, c_int, c_int, _P]),
    "rpb200_halo_exchange_unpack_seg": (c_int, [_P, c_int, c_int, c_int, _P]),
    "rpb200_halo_sendrecv_bind": (c_int, [_P, POINTER(_P)]),
    "rpb200_halo_sendrecv": (c_int, [_P, _P]),
    "rpb200_halo_sendrecv_put": (c_int, [_P, _P]),
    "rpb200_halo_sendrecv_wait": (c_int, [_P, _P]),
    "rpb200_halo_recv_buffer": (c_int, [_P, c_int, POINTER(_P), POINTER(c_int64)]),
    "rpb200_halo_exchange": (c_int, [_P, _P]),
    "rpb200_halo_exchange_status": (c_int, [_P]),
    "rpb200_ipc_export": (c_int, [_P, POINTER(c_ubyte)]),
    "rpb200_ipc_open": (c_int, [POINTER(c_ubyte), POINTER(_P)]),
    "rpb200_ipc_close": (c_int, [_P]),
    "rpb200_enable_peer_access": (c_int, [c_int, c_int]),
    "rpb200_malloc": (c_int, [POINTER(_P), c_size_t]),
    "rpb200_free": (c_int, [_P]),
    "rpb200_malloc_host": (c_int, [POINTER(_P), c_size_t]),
    "rpb200_free_host": (c_int, [_P]),
    "rpb200_memcpy_h2d": (c_int, [_P, _P, c_size_t, _P]),
    "rpb200_memcpy_d2h": (c_int, [_P, _P, c_size_t, _P]),
    "rpb200_memset": (c_int, [_P, c_int, c_size_t, _P]),
    "rpb200_stream_synchronize": (c_int, [_P]),
    "rpb200_device_synchronize": (c_int, []),
    "rpb200_timer_create": (c_int, [POINTER(_P)]),
    "rpb200_timer_start": (c_int, [_P, _P]),
    "rpb200_timer_stop": (c_int, [_P, _P]),
    "rpb200_timer_elapsed_ms": (c_int, [_P, POINTER(c_float)]),
    "rpb200_timer_destroy": (None, [_P]),
}

_lib = None


def load():
    """Load librpb200.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RPB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the Base_B200 kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(err: int, what: str = "rpb200 call"):
    if err != 0:
        msg = load().rpb200_error_string(err)
        raise RPB200Error(f"{what} failed: {err} ({msg.decode() if msg else '?'})")


def _ptr(t):
    """Device (or pinned host) pointer of a torch tensor / raw int / None."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


class Context:
    """Owns one rpb200_ctx.  Methods mirror the C entry points 1:1 and enqueue on torch's
    current CUDA stream; tensors must be contiguous float64 CUDA tensors on this device."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = c_void_p()
        check(self.lib.rpb200_create(device, ctypes.byref(h)), "rpb200_create")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.rpb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self) -> int:
        return self.lib.rpb200_sm_count(self.h)

    def set_tuning(self, kernel: str, block_size: int = -1, ctas_per_sm: int = -1, unroll: int = -1):
        check(self.lib.rpb200_set_tuning(self.h, kernel.encode(), block_size, ctas_per_sm, unroll),
              f"set_tuning({kernel})")

    def get_tuning(self, kernel: str):
        """(block_size, ctas_per_sm, unroll) currently in force for `kernel`."""
        b, c, u = c_int(), c_int(), c_int()
        check(self.lib.rpb200_get_tuning(self.h, kernel.encode(), ctypes.byref(b), ctypes.byref(c), ctypes.byref(u)),
              f"get_tuning({kernel})")
        return b.value, c.value, u.value

    def reset_tuning(self, kernel: str | None = None):
        """Built-in launch shape of `kernel` (None: of every kernel) again."""
        check(self.lib.rpb200_reset_tuning(self.h, kernel.encode() if kernel else None), f"reset_tuning({kernel})")

    def stream_attach(self, stream=None):
        """Give a stream (default: torch's current one) its scratch set now (rpb200.h: needed before capturing on the 5th+
        stream of a context)."""
        check(self.lib.rpb200_stream_attach(self.h, _stream() if stream is None else stream), "stream_attach")

    def stream_detach(self, stream=None):
        check(self.lib.rpb200_stream_detach(self.h, _stream() if stream is None else stream), "stream_detach")

    def scan_reserve(self, n: int):
        check(self.lib.rpb200_scan_reserve(self.h, n), "scan_reserve")

    def indexlist_reserve(self, n: int):
        check(self.lib.rpb200_indexlist_reserve(self.h, n), "indexlist_reserve")

    # ---- Stream --------------------------------------------------------------------------
    def stream_copy(self, c, a, n=None):
        n = a.numel() if n is None else n
        check(self.lib.rpb200_stream_copy(self.h, _ptr(c), _ptr(a), n, _stream()), "stream_copy")

    def stream_mul(self, b, c, alpha, n=None):
        n = c.numel() if n is None else n
        check(self.lib.rpb200_stream_mul(self.h, _ptr(b), _ptr(c), alpha, n, _stream()), "stream_mul")

    def stream_add(self, c, a, b, n=None):
        n = a.numel() if n is None else n
        check(self.lib.rpb200_stream_add(self.h, _ptr(c), _ptr(a), _ptr(b), n, _stream()), "stream_add")

    def stream_triad(self, a, b, c, alpha, n=None):
        n = b.numel() if n is None else n
        check(self.lib.rpb200_stream_triad(self.h, _ptr(a), _ptr(b), _ptr(c), alpha, n, _stream()),
              "stream_triad")

    def stream_dot(self, a, b, out, init=0.0, accumulate=False, n=None):
        n = a.numel() if n is None else n
        check(self.lib.rpb200_stream_dot(self.h, _ptr(a), _ptr(b), n, init, _ptr(out),
                                         1 if accumulate else 0, _stream()), "stream_dot")

    # ---- Algorithm -----------------------------------------------------------------------
    def reduce_sum(self, x, out, init=0.0, n=None):
        n = x.numel() if n is None else n
        check(self.lib.rpb200_reduce_sum(self.h, _ptr(x), n, init, _ptr(out), _stream()), "reduce_sum")

    def scan_exclusive(self, x, y, n=None):
        n = x.numel() if n is None else n
        check(self.lib.rpb200_scan_exclusive(self.h, _ptr(x), _ptr(y), n, _stream()), "scan_exclusive")

    def sort_scratch_bytes(self, n: int, pairs: bool = False) -> int:
        return self.lib.rpb200_sort_scratch_bytes(n, 1 if pairs else 0)

    def sort_keys(self, keys, scratch, n=None):
        n = keys.numel() if n is None else n
        check(self.lib.rpb200_sort_keys_f64(self.h, _ptr(keys), n, _ptr(scratch),
                                            scratch.numel() * scratch.element_size(), _stream()),
              "sort_keys_f64")

    def sort_pairs(self, keys, vals, scratch, n=None):
        n = keys.numel() if n is None else n
        check(self.lib.rpb200_sort_pairs_f64(self.h, _ptr(keys), _ptr(vals), n, _ptr(scratch),
                                             scratch.numel() * scratch.element_size(), _stream()),
              "sort_pairs_f64")

    # ---- Apps ----------------------------------------------------------------------------
    def mass3dpa(self, B, Bt, D, X, Y, NE):
        check(self.lib.rpb200_mass3dpa(self.h, _ptr(B), _ptr(Bt), _ptr(D), _ptr(X), _ptr(Y), NE,
                                       _stream()), "mass3dpa")

    def diffusion3dpa(self, B, G, D, X, Y, NE, symmetric=True):
        check(self.lib.rpb200_diffusion3dpa(self.h, _ptr(B), _ptr(G), _ptr(D), _ptr(X), _ptr(Y), NE,
                                            1 if symmetric else 0, _stream()), "diffusion3dpa")

    def convection3dpa(self, B, Bt, G, D, X, Y, NE):
        check(self.lib.rpb200_convection3dpa(self.h, _ptr(B), _ptr(Bt), _ptr(G), _ptr(D), _ptr(X),
                                             _ptr(Y), NE, _stream()), "convection3dpa")

    def ltimes(self, phi, ell, psi, num_d, num_g, num_m, num_z):
        check(self.lib.rpb200_ltimes(self.h, _ptr(phi), _ptr(ell), _ptr(psi), num_d, num_g, num_m,
                                     num_z, _stream()), "ltimes")

    # ---- widened rows -----------------------------------------------------
    def indexlist(self, x, list_, d_len, n=None):
        """list_: int32 tensor (>= n entries), d_len: int64 tensor of 1 element (device)."""
        n = x.numel() if n is None else n
        check(self.lib.rpb200_indexlist(self.h, _ptr(x), _ptr(list_), n, _ptr(d_len), _stream()), "indexlist")

    def memset_f64(self, x, val, n=None):
        n = x.numel() if n is None else n
        check(self.lib.rpb200_memset_f64(self.h, _ptr(x), val, n, _stream()), "memset_f64")

    def polybench_gemm(self, A, B, C, ni, nj, nk, alpha, beta=0.0):
        check(self.lib.rpb200_polybench_gemm(self.h, _ptr(A), _ptr(B), _ptr(C), ni, nj, nk, alpha, beta,
                                             _stream()), "polybench_gemm")

    # ---- Comm ----------------------------------------------------------------------------
    def halo_chunk(self) -> int:
        return self.lib.rpb200_halo_chunk()

    def halo_worklist(self, segs):
        """segs: iterable of (buffer_ptr, list_ptr, var_ptr, len, msg) -> HaloWorklist."""
        return HaloWorklist(self, segs)

    def halo_pack(self, wl):
        check(self.lib.rpb200_halo_pack(self.h, wl.h, _stream()), "halo_pack")

    def halo_unpack(self, wl):
        check(self.lib.rpb200_halo_unpack(self.h, wl.h, _stream()), "halo_unpack")

    def halo_pack_unpack(self, pack_wl, unpack_wl):
        """pack + unpack of one rep in ONE launch (the two lists must touch disjoint memory: rpb200.h)."""
        check(self.lib.rpb200_halo_pack_unpack(self.h, pack_wl.h, unpack_wl.h, _stream()), "halo_pack_unpack")

    def halo_plan(self, grid_dims, halo_width=1, num_vars=3, rank=0, rank_dims=(1, 1, 1)):
        return HaloPlan(self, grid_dims, halo_width, num_vars, rank, rank_dims)


def halo_grid_dims(target_size: int):
    """HALO_base.cpp:31-35."""
    d = (c_int64 * 3)()
    load().rpb200_halo_grid_dims(target_size, d)
    return [int(x) for x in d]


def _seg_array(segs):
    arr = (HaloSeg * max(len(segs), 1))()
    for i, (buf, lst, var, ln, msg) in enumerate(segs):
        arr[i] = HaloSeg(_ptr(buf), _ptr(lst), _ptr(var), ln, msg, 0)
    return arr


class HaloWorklist:
    """rpb200_halo_worklist: the reference's (buffer, list, var, len) tuples, device-resident."""

    def __init__(self, ctx, segs):
        self.ctx = ctx
        segs = list(segs)
        self.n = len(segs)
        h = c_void_p()
        check(ctx.lib.rpb200_halo_worklist_create(ctx.h, ctypes.cast(_seg_array(segs), c_void_p), self.n,
                                                  ctypes.byref(h)), "halo_worklist_create")
        self.h = h

    def update(self, segs):
        segs = list(segs)
        check(self.ctx.lib.rpb200_halo_worklist_update(self.h, ctypes.cast(_seg_array(segs), c_void_p), len(segs),
                                                       _stream()), "halo_worklist_update")

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.rpb200_halo_worklist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HaloPlan:
    """rpb200_halo_plan: HALO_base (index lists + rank grid) and the fused pack/unpack/exchange."""

    def __init__(self, ctx, grid_dims, halo_width, num_vars, rank, rank_dims):
        self.ctx, self.lib = ctx, ctx.lib
        self.num_vars, self.rank = num_vars, rank
        self.nranks = rank_dims[0] * rank_dims[1] * rank_dims[2]
        gd = (c_int64 * 3)(*grid_dims)
        rd = (c_int * 3)(*rank_dims)
        h = c_void_p()
        check(self.lib.rpb200_halo_plan_create(ctx.h, gd, halo_width, num_vars, rank, rd, ctypes.byref(h)),
              "halo_plan_create")
        self.h = h
        self.var_size = self.lib.rpb200_halo_plan_var_size(h)
        self.neighbors = []
        for l in range(26):
            r, st, rt = c_int(), c_int(), c_int()
            pl, ul = c_int64(), c_int64()
            dp, du = c_void_p(), c_void_p()
            check(self.lib.rpb200_halo_plan_neighbor(h, l, ctypes.byref(r), ctypes.byref(st), ctypes.byref(rt),
                                                     ctypes.byref(pl), ctypes.byref(ul), ctypes.byref(dp),
                                                     ctypes.byref(du)), "halo_plan_neighbor")
            self.neighbors.append(dict(rank=r.value, send_tag=st.value, recv_tag=rt.value, pack_len=pl.value,
                                       unpack_len=ul.value, d_pack_list=dp.value, d_unpack_list=du.value))
        self._keep = None

    @staticmethod
    def _ptr_array(items):
        arr = (c_void_p * len(items))()
        for i, t in enumerate(items):
            arr[i] = _ptr(t)
        return arr

    def bind(self, vars_, pack_buffers, unpack_buffers):
        self._keep = (vars_, pack_buffers, unpack_buffers)
        check(self.lib.rpb200_halo_plan_bind(self.h, self._ptr_array(vars_), self._ptr_array(pack_buffers),
                                             self._ptr_array(unpack_buffers)), "halo_plan_bind")

    def pack(self):
        check(self.lib.rpb200_halo_plan_pack(self.h, _stream()), "halo_plan_pack")

    def unpack(self):
        check(self.lib.rpb200_halo_plan_unpack(self.h, _stream()), "halo_plan_unpack")

    def pack_unpack(self):
        """One rep of HALO_PACKING_FUSED in ONE launch (pack and unpack items interleaved)."""
        check(self.lib.rpb200_halo_plan_pack_unpack(self.h, _stream()), "halo_plan_pack_unpack")

    def window(self, vars_, want_handle=True):
        """Allocate this rank's receive window; returns (device pointer, bytes, ipc handle bytes).
        vars_ = None: a transport-only window (HALO_SENDRECV)."""
        self._keep_x = vars_
        w, nb = c_void_p(), c_size_t()
        hbuf = (c_ubyte * 64)()
        check(self.lib.rpb200_halo_exchange_window(self.h, self._ptr_array(vars_) if vars_ is not None else None, ctypes.byref(w), ctypes.byref(nb),
                                                   hbuf if want_handle else None), "halo_exchange_window")
        return w.value, nb.value, bytes(hbuf)

    def connect(self, handles):
        """handles: list of nranks 64-byte IPC handles, indexed by rank."""
        blob = b"".join(handles)
        buf = (c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self.lib.rpb200_halo_exchange_connect(self.h, len(handles), buf), "halo_exchange_connect")

    def connect_ptrs(self, windows):
        arr = (c_void_p * len(windows))(*windows)
        check(self.lib.rpb200_halo_exchange_connect_ptrs(self.h, len(windows), arr), "halo_exchange_connect_ptrs")

    def exchange_pack(self):
        check(self.lib.rpb200_halo_exchange_pack(self.h, _stream()), "halo_exchange_pack")

    def exchange_unpack(self):
        check(self.lib.rpb200_halo_exchange_unpack(self.h, _stream()), "halo_exchange_unpack")

    def exchange(self):
        check(self.lib.rpb200_halo_exchange(self.h, _stream()), "halo_exchange")

    def sendrecv_bind(self, send_buffers):
        self._keep_send = send_buffers
        check(self.lib.rpb200_halo_sendrecv_bind(self.h, self._ptr_array(send_buffers)), "halo_sendrecv_bind")

    def sendrecv(self):
        check(self.lib.rpb200_halo_sendrecv(self.h, _stream()), "halo_sendrecv")

    def sendrecv_put(self):
        check(self.lib.rpb200_halo_sendrecv_put(self.h, _stream()), "halo_sendrecv_put")

    def sendrecv_wait(self):
        check(self.lib.rpb200_halo_sendrecv_wait(self.h, _stream()), "halo_sendrecv_wait")

    def recv_buffer(self, l):
        """(device pointer, length in doubles) of message l of the last completed rep."""
        p, n = c_void_p(), c_int64()
        check(self.lib.rpb200_halo_recv_buffer(self.h, l, ctypes.byref(p), ctypes.byref(n)), "halo_recv_buffer")
        return p.value, n.value

    def exchange_unfused(self, num_vars):
        """The unfused HALO_EXCHANGE: one pack launch and one unpack launch per (neighbour, variable)."""
        st = _stream()
        for l in range(26):
            for v in range(num_vars):
                check(self.lib.rpb200_halo_exchange_pack_seg(self.h, l, v, st), "halo_exchange_pack_seg")
        for l in range(26):
            for v in range(num_vars):
                last = 1 if (l == 25 and v == num_vars - 1) else 0
                check(self.lib.rpb200_halo_exchange_unpack_seg(self.h, l, v, last, st), "halo_exchange_unpack_seg")

    def status(self):
        check(self.lib.rpb200_halo_exchange_status(self.h), "halo_exchange_status")

    def close(self):
        if getattr(self, "h", None):
            self.lib.rpb200_halo_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ipc_export(d_ptr: int) -> bytes:
    lib = load()
    buf = (c_ubyte * 64)()
    check(lib.rpb200_ipc_export(d_ptr, buf), "ipc_export")
    return bytes(buf)


def ipc_open(handle: bytes) -> int:
    lib = load()
    buf = (c_ubyte * 64).from_buffer_copy(handle)
    out = c_void_p()
    check(lib.rpb200_ipc_open(buf, ctypes.byref(out)), "ipc_open")
    return out.value


def ipc_close(d_ptr: int):
    check(load().rpb200_ipc_close(d_ptr), "ipc_close")

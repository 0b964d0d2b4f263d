"""Suite-generated inputs for each kernel, produced by the oracle's restatement of the
reference setUp() sequences.  Test infrastructure only."""
from __future__ import annotations

import numpy as np

import oracle


def _seq(calls):
    """Run a setUp-like sequence of init calls after a counter reset.
    calls: list of ('real', n) | ('const', n, v) | ('rand', n) | ('scalar',)"""
    L = oracle.lib()
    L.orc_reset_init_count()
    out = []
    for c in calls:
        if c[0] == "real":
            a = np.empty(c[1]); L.orc_init_real(a, c[1]); out.append(a)
        elif c[0] == "const":
            a = np.empty(c[1]); L.orc_init_const(a, c[1], c[2]); out.append(a)
        elif c[0] == "rand":
            a = np.empty(c[1]); L.orc_init_rand_value(a, c[1]); out.append(a)
        elif c[0] == "scalar":
            out.append(L.orc_init_scalar())
    return out


def stream_copy(n):   # stream/COPY.cpp:71-72
    a, c = _seq([("real", n), ("const", n, 0.0)])
    return dict(a=a, c=c)


def stream_mul(n):    # stream/MUL.cpp:71-73
    b, c, alpha = _seq([("const", n, 0.0), ("real", n), ("scalar",)])
    return dict(b=b, c=c, alpha=alpha)


def stream_add(n):    # stream/ADD.cpp:71-73
    a, b, c = _seq([("real", n), ("real", n), ("const", n, 0.0)])
    return dict(a=a, b=b, c=c)


def stream_triad(n):  # stream/TRIAD.cpp:75-78
    a, b, c, alpha = _seq([("const", n, 0.0), ("real", n), ("real", n), ("scalar",)])
    return dict(a=a, b=b, c=c, alpha=alpha)


def stream_dot(n):    # stream/DOT.cpp:64-69
    a, b = _seq([("real", n), ("real", n)])
    return dict(a=a, b=b)


def reduce_sum(n):    # algorithm/REDUCE_SUM.cpp:62-66
    (x,) = _seq([("real", n)])
    return dict(x=x)


def scan(n):          # algorithm/SCAN.cpp:68-71
    x, y = _seq([("rand", n), ("const", n, 0.0)])
    return dict(x=x, y=y)


def indexlist(n):     # basic/INDEXLIST.cpp:62-67 (and INDEXLIST_3LOOP.cpp: same setUp)
    L = oracle.lib()
    L.orc_reset_init_count()
    x = np.empty(n); L.orc_init_rand_sign(x, n)
    lst = np.empty(max(n, 1), dtype=np.int32); L.orc_init_int(lst, n)
    return dict(x=x, list=lst[:n] if n else lst[:0])


def polybench_gemm(target=0):   # polybench/POLYBENCH_GEMM.cpp:21-89
    L = oracle.lib()
    ni, nj, nk = (np.zeros(1, dtype=np.int64) for _ in range(3))
    L.orc_polybench_gemm_dims(target, ni, nj, nk)
    ni, nj, nk = int(ni[0]), int(nj[0]), int(nk[0])
    A, B, C = _seq([("real", ni * nk), ("real", nk * nj), ("const", ni * nj, 0.0)])
    scale = float(np.longdouble(0.001) * (np.longdouble(1000 * 1000) / np.longdouble(ni * nj)))
    return dict(A=A, B=B, C=C, ni=ni, nj=nj, nk=nk, alpha=0.62, beta=1.002, scale=scale)


def triad_scale(n):   # stream/TRIAD.cpp:36-38 (long double, narrowed to double at the call)
    return float(np.longdouble(0.001) * (np.longdouble(1000000) / np.longdouble(n)))


def scan_scale(n):    # algorithm/SCAN.cpp:36-39
    return float(np.longdouble(1e-2) * (np.longdouble(1000000) / np.longdouble(n)) / np.longdouble(n))


def round_div(target, unit):
    return max((target + unit // 2) // unit, 1)


def mass3dpa(target=0):   # apps/MASS3DPA.cpp:23-83
    NE = round_div(target or 8000 * 125, 125)
    return dict(NE=NE, B=np.ones(20), Bt=np.ones(20), D=np.ones(125 * NE), X=np.ones(64 * NE),
                Y=np.zeros(64 * NE))


def diffusion3dpa(target=0):   # apps/DIFFUSION3DPA.cpp:23-88
    NE = round_div(target or 15625 * 64, 64)
    return dict(NE=NE, B=np.ones(12), G=np.ones(12), D=np.ones(64 * 6 * NE), X=np.ones(27 * NE),
                Y=np.zeros(27 * NE))


def convection3dpa(target=0):   # apps/CONVECTION3DPA.cpp:23-89
    NE = round_div(target or 15625 * 64, 64)
    return dict(NE=NE, B=np.ones(12), Bt=np.ones(12), G=np.ones(12), D=np.ones(64 * 3 * NE),
                X=np.ones(27 * NE), Y=np.zeros(27 * NE))


def ltimes(target=0, nd=64, ng=32, nm=25):   # apps/LTIMES.cpp:23-92
    dg = nd * ng
    dflt = dg * round_div(1000000, dg)
    nz = round_div(target or dflt, dg)
    phi, ell, psi = _seq([("const", nm * ng * nz, 0.0), ("real", nd * nm), ("real", dg * nz)])
    scale = float(np.longdouble(0.001) * (np.longdouble(dflt) / np.longdouble(dg * nz)))
    return dict(nz=nz, nd=nd, ng=ng, nm=nm, phi=phi, ell=ell, psi=psi, scale=scale)


def halo_lists(dims, hw):
    """The 52 index lists of HALO_base::create_lists (comm/HALO_base.cpp:169-291) from the oracle."""
    L = oracle.lib()
    d = np.asarray(dims, dtype=np.int64)
    pack, unpack = [], []
    for l in range(26):
        for recv, out in ((0, pack), (1, unpack)):
            n = L.orc_halo_extent_len(recv, l, hw, d)
            lst = np.empty(n, dtype=np.int32)
            L.orc_halo_make_list(recv, l, hw, d, lst)
            out.append(lst)
    return pack, unpack


def halo_neighbors(rank, pdims):
    L = oracle.lib()
    r, st, rt = (np.empty(26, dtype=np.int32) for _ in range(3))
    L.orc_halo_neighbors(rank, np.asarray(pdims, dtype=np.int32), r, st, rt)
    return r, st, rt


def halo_packing_fused(target=0, hw=1, nvars=3):
    """setUp of Comm_HALO_PACKING_FUSED (HALO_base.cpp:52-67, HALO_PACKING_FUSED.cpp:63-109): the
    init counter is bumped by the 52 list allocations and the vars, so pack buffer l is filled
    with initData at count 52+nvars+l and unpack buffer l at 52+nvars+26+l."""
    L = oracle.lib()
    dims = np.zeros(3, dtype=np.int64)
    L.orc_halo_grid_dims(target or 1000000, dims)
    pack, unpack = halo_lists(dims, hw)
    var_size = int(np.prod(dims + 2 * hw))
    L.orc_reset_init_count()
    dummy = np.zeros(1, dtype=np.int32)
    for _ in range(52):
        L.orc_init_int(dummy, 0)
    vars_ = []
    for v in range(nvars):
        a = np.empty(var_size); L.orc_init_real(a, 0)
        vars_.append(np.arange(var_size, dtype=np.float64) + v)
    pack_bufs, unpack_bufs = [], []
    for l in range(26):
        a = np.empty(nvars * pack[l].size); L.orc_init_real(a, a.size); pack_bufs.append(a)
    for l in range(26):
        a = np.empty(nvars * unpack[l].size); L.orc_init_real(a, a.size); unpack_bufs.append(a)
    return dict(dims=[int(x) for x in dims], hw=hw, nvars=nvars, var_size=var_size, pack_lists=pack,
                unpack_lists=unpack, vars=vars_, pack_bufs=pack_bufs, unpack_bufs=unpack_bufs)

#!/usr/bin/env python
"""Launch each hot-path kernel ONCE at a profiling size, for `ncu --set full` (run under gpurun):

    ncu --set full --clock-control none --import-source on -k regex:'stream_ew|reduce_kernel|scan_kernel|sort_|pa_kernel|ltimes|halo_kernel' \
        -o gpurun_out/r01_prof python tools/prof_kernels.py [which ...]

Sizes are the BASELINE sizes for the streaming kernels and 1/4 of them for the PA kernels (ncu saves and
restores every written buffer around each of its ~40 replay passes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402

which = set(sys.argv[1:]) or {"stream", "scan", "sort", "pa", "ltimes", "halo", "gemm", "indexlist"}
ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")

if "stream" in which:
    n = 1 << 28
    a = torch.rand(n, **f64); b = torch.rand(n, **f64); c = torch.empty(n, **f64)
    out = torch.zeros(1, **f64)
    ctx.stream_copy(c, a); ctx.stream_mul(c, a, 0.3); ctx.stream_add(c, a, b); ctx.stream_triad(c, a, b, 0.3)
    ctx.stream_dot(a, b, out); ctx.reduce_sum(a, out, n=1 << 27)
    torch.cuda.synchronize(); del a, b, c
if "scan" in which:
    n = 1 << 27
    x = torch.rand(n, **f64); y = torch.empty(n, **f64)
    ctx.scan_exclusive(x, y)
    torch.cuda.synchronize(); del x, y
if "sort" in which:
    n = 1 << 27
    x = torch.rand(n, **f64)
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 32, **f64)
    ctx.sort_keys(x, scratch)
    if "sortpairs" in which:
        v = torch.rand(n, **f64)
        ctx.sort_pairs(x, v, scratch)
    torch.cuda.synchronize(); del x, scratch
if "pa" in which:
    NE = 1000000
    one = lambda m: torch.ones(m, **f64)
    B, Bt, D, X, Y = one(20), one(20), one(125 * NE), one(64 * NE), one(64 * NE)
    ctx.mass3dpa(B, Bt, D, X, Y, NE)
    torch.cuda.synchronize(); del D, X, Y
    B, G, D, X, Y = one(12), one(12), one(384 * NE), one(27 * NE), one(27 * NE)
    ctx.diffusion3dpa(B, G, D, X, Y, NE)
    torch.cuda.synchronize(); del D
    D = one(192 * NE)
    ctx.convection3dpa(B, B, G, D, X, Y, NE)
    torch.cuda.synchronize(); del D, X, Y
if "ltimes" in which:
    nz = 125000
    phi = torch.zeros(800 * nz, **f64); psi = torch.rand(2048 * nz, **f64); ell = torch.rand(1600, **f64)
    ctx.ltimes(phi, ell, psi, 64, 32, 25, nz)
    torch.cuda.synchronize(); del phi, psi
if "halo" in which:
    g, nv = 512, 3
    plan = ctx.halo_plan((g, g, g), 1, nv)
    vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
    pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
    ub = [torch.zeros(nv * nb["unpack_len"], **f64) for nb in plan.neighbors]
    plan.bind(vars_, pb, ub)
    plan.pack(); plan.unpack()
    plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
    plan.exchange()
    torch.cuda.synchronize()
    plan.status()
if "gemm" in which:
    ni = nj = 4096; nk = 4915
    A = torch.rand(ni * nk, **f64); B = torch.rand(nk * nj, **f64); C = torch.empty(ni * nj, **f64)
    ctx.polybench_gemm(A, B, C, ni, nj, nk, 0.62)
    torch.cuda.synchronize()
if "indexlist" in which:
    n = 1 << 27
    x = torch.randn(n, **f64); lst = torch.empty(n, dtype=torch.int32, device="cuda"); ln = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.indexlist(x, lst, ln)
    torch.cuda.synchronize()
print("prof_kernels done")

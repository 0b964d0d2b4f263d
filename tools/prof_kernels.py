#!/usr/bin/env python
"""Launch each hot-path kernel ONCE at a profiling size, for `ncu --set full` (run under gpurun):

    ncu --set full --clock-control none --import-source on -k regex:'stream_ew|reduce_kernel|scan_kernel|sort_|pa_kernel|ltimes|halo_kernel' \
        -o gpurun_out/r01_prof python tools/prof_kernels.py [which ...]

Sizes are the BASELINE sizes for every kernel.  `--manifest FILE` writes, per suite kernel, the size and the launch
shape (rpb200_get_tuning) this process launched: tools/ncu_traffic.py joins it with the ncu raw CSV into
profiles/r02_ncu_traffic.json, which bench.py only trusts for a run of the same size and launch shape."""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402

argv = sys.argv[1:]
manifest_path = None
if "--manifest" in argv:
    manifest_path = argv[argv.index("--manifest") + 1]
    argv = [a for a in argv if a not in ("--manifest", manifest_path)]
which = set(argv) or {"stream", "scan", "sort", "pa", "ltimes", "halo", "gemm", "indexlist"}
ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")
manifest = {}


def note(kernel, n, **kw):
    manifest[kernel] = dict(n=n, tuning=list(ctx.get_tuning(kernel)), **kw)


if "stream" in which:
    n = 1 << 28
    a = torch.rand(n, **f64); b = torch.rand(n, **f64); c = torch.empty(n, **f64)
    out = torch.zeros(1, **f64)
    ctx.stream_copy(c, a); ctx.stream_mul(c, a, 0.3); ctx.stream_add(c, a, b); ctx.stream_triad(c, a, b, 0.3)
    ctx.stream_dot(a, b, out); ctx.reduce_sum(a, out, n=1 << 27)
    for k in ("Stream_COPY", "Stream_MUL", "Stream_ADD", "Stream_TRIAD", "Stream_DOT"):
        note(k, n)
    note("Algorithm_REDUCE_SUM", 1 << 27)
    torch.cuda.synchronize(); del a, b, c
if "scan" in which:
    n = 1 << 27
    x = torch.rand(n, **f64); y = torch.empty(n, **f64)
    ctx.scan_exclusive(x, y)
    note("Algorithm_SCAN", n)
    torch.cuda.synchronize(); del x, y
if "sort" in which:
    n = 1 << 27
    x = torch.randint(0, 2**31 - 1, (n,), device="cuda").to(torch.float64).div_(2147483647.0)      # rand()/RAND_MAX
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 32, **f64)
    ctx.sort_keys(x, scratch)
    note("Algorithm_SORT", n)
    if "sortpairs" in which:
        v = torch.rand(n, **f64)
        ctx.sort_pairs(x, v, scratch)
    torch.cuda.synchronize(); del x, scratch
if "pa" in which:
    NE = 4000000
    one = lambda m: torch.ones(m, **f64)
    for k in ("Apps_MASS3DPA", "Apps_DIFFUSION3DPA", "Apps_CONVECTION3DPA"):
        note(k, NE)
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
    nz = 500000
    note("Apps_LTIMES", 32 * nz)
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
    ne = sum(nb["pack_len"] for nb in plan.neighbors) * nv
    note("Comm_HALO_PACKING_FUSED", ne); note("Comm_HALO_EXCHANGE_FUSED", ne)
    plan.pack_unpack()
    plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
    plan.exchange()
    torch.cuda.synchronize()
    plan.status()
if "gemm" in which:
    ni = nj = 4096; nk = 4915
    A = torch.rand(ni * nk, **f64); B = torch.rand(nk * nj, **f64); C = torch.empty(ni * nj, **f64)
    ctx.polybench_gemm(A, B, C, ni, nj, nk, 0.62)
    note("Polybench_GEMM", ni * nj, dims=[ni, nj, nk])
    torch.cuda.synchronize()
if "indexlist" in which:
    n = 1 << 27
    x = torch.randn(n, **f64); lst = torch.empty(n, dtype=torch.int32, device="cuda"); ln = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.indexlist(x, lst, ln)
    note("Basic_INDEXLIST", n)
    torch.cuda.synchronize()
if manifest_path:
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rajaperf_b200", "lib", "librpb200.so")
    json.dump({"library_sha256": hashlib.sha256(open(so, "rb").read()).hexdigest(), "kernels": manifest}, open(manifest_path, "w"), indent=1)
print("prof_kernels done")

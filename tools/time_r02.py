#!/usr/bin/env python
"""Round-2 A/B timings on one B200 (run under gpurun):  python tools/time_r02.py halo halo1024 sort reduce [--out f.json]

  halo      HALO_PACKING_FUSED 512^3 x 3 vars: the two-launch form (round 1) against the one-launch item-list form (x-face
            items mixed in / first), CTAs per SM, and the L2 fetch granularity (cudaLimitMaxL2FetchGranularity 64 / 32 / 128);
            HALO_EXCHANGE_FUSED on the 1 x 1 x 1 rank grid: two launches against one
  halo1024  the same at 1024^3 (SURVEY 8d's second halo configuration, --size 1073741824)
  sort      SORT / SORTPAIRS at 2^27 rand()/RAND_MAX keys: lane-private against shared-bin histograms, torch.sort for scale
  reduce    DOT (2^28) / REDUCE_SUM (2^27) launch-shape sweep
Every timing: CUDA events around 100 reps captured in one CUDA graph (halo) or around the rep loop (others), after warm-up.
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
out_path = None
if "--out" in sys.argv:
    out_path = sys.argv[sys.argv.index("--out") + 1]
    args = [a for a in args if a != out_path]
which = set(args) or {"halo", "sort", "reduce"}
ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")
res = {}


def time_ms(fn, reps=20, warm=3, setup=None):
    for _ in range(warm):
        if setup: setup()
        fn()
    torch.cuda.synchronize()
    if setup is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    tot = 0.0
    for _ in range(reps):
        setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def graph_ms(body, reps=100):
    body(); torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(reps):
            body()
    g_.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


def report(name, bytes_, ms, **kw):
    res[name] = dict(ms=ms, gbs=bytes_ / ms / 1e6, **kw)
    print(f"{name:72s} {ms * 1e3:9.2f} us {bytes_ / ms / 1e6:9.1f} GB/s {kw if kw else ''}", flush=True)


def set_l2_fetch(nbytes):
    """cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity = 5, nbytes): a hint, device-wide (same primary context as the
    library's statically linked runtime)."""
    rt = ctypes.CDLL("libcudart.so.12")
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(nbytes))
    got = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
    return rc, got.value


def halo(g):
    nv = 3
    plan = ctx.halo_plan((g, g, g), 1, nv)
    vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
    pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
    ub = [torch.zeros(nv * nb["unpack_len"], **f64) for nb in plan.neighbors]
    plan.bind(vars_, pb, ub)
    ne = sum(nb["pack_len"] for nb in plan.neighbors) * nv
    plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
    K = "Comm_HALO_PACKING_FUSED"
    for rnd in range(2):
        ctx.set_tuning(K, 192, 4, 2)
        report(f"halo{g} pack+unpack TWO launches (r01 default) round {rnd}", 40 * ne, graph_ms(lambda: (plan.pack(), plan.unpack())))
        for cps, order, label in ((4, 1, "x units mixed in"), (2, 1, "x units mixed in"), (1, 1, "x units mixed in"),
                                  (2, 3, "x units first"), (2, 5, "two phases")):
            ctx.set_tuning(K, 192, cps, order)
            report(f"halo{g} pack+unpack ONE launch, {label}, {cps} CTAs/SM round {rnd}", 40 * ne, graph_ms(plan.pack_unpack))
    ctx.reset_tuning(K)
    X = "Comm_HALO_EXCHANGE_FUSED"
    for rnd in range(2):
        for cps in (4, 2, 1):
            ctx.set_tuning(X, 192, cps, 2)
            report(f"halo{g} exchange 1 rank, TWO launches (r01 default), {cps} CTAs/SM round {rnd}", 56 * ne, graph_ms(plan.exchange))
            ctx.set_tuning(X, 192, cps, 1)
            report(f"halo{g} exchange 1 rank, ONE launch (item list), {cps} CTAs/SM round {rnd}", 56 * ne, graph_ms(plan.exchange))
    ctx.reset_tuning(X)
    plan.status()
    plan.close()


if "halo" in which:
    halo(512)
if "halo1024" in which:
    halo(1024)

if "sort" in which:
    n = 1 << 27
    src = torch.randint(0, 2**31 - 1, (n,), device="cuda").to(torch.float64).div_(2147483647.0)      # rand()/RAND_MAX
    k = torch.empty_like(src); v = torch.empty_like(src)
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 32, **f64)
    for rnd in range(2):
        for var, lb, label in ((4, 1, "look-back one tile at a time"), (4, 2, "look-back two tiles at a time"),
                               (4, 4, "default: lane-private histogram, uniform-tile test, look-back 4 at a time")):
            ctx.set_tuning("Algorithm_SORT", -1, lb, var); ctx.set_tuning("Algorithm_SORTPAIRS", -1, lb, var)
            ms = time_ms(lambda: ctx.sort_keys(k, scratch), 5, 2, setup=lambda: k.copy_(src))
            report(f"sort keys 2^27, {label} round {rnd}", 16 * n, ms, mkeys_s=n / ms / 1e3)
            ms = time_ms(lambda: ctx.sort_pairs(k, v, scratch), 5, 2, setup=lambda: (k.copy_(src), v.copy_(src)))
            report(f"sort pairs 2^27, {label} round {rnd}", 32 * n, ms, mkeys_s=n / ms / 1e3)
    ctx.reset_tuning("Algorithm_SORT"); ctx.reset_tuning("Algorithm_SORTPAIRS")
    ok = bool((k[1:] >= k[:-1]).all())
    print("sorted:", ok, flush=True)
    res["sort sorted"] = ok
    ms = time_ms(lambda: torch.sort(src), 5, 2)
    report("torch.sort (CUB pairs: keys + int64 indices)", 32 * n, ms, mkeys_s=n / ms / 1e3)
    # uniform keys (every tile uniform in every pass): the floor of the pass structure
    k.fill_(0.37)
    ms = time_ms(lambda: ctx.sort_keys(k, scratch), 5, 2)
    report("sort keys 2^27, all keys equal (every tile uniform)", 16 * n, ms, mkeys_s=n / ms / 1e3)
    del src, k, v, scratch

if "reduce" in which:
    n = 1 << 28
    a = torch.rand(n, **f64); b = torch.rand(n, **f64); o = torch.zeros(1, **f64)
    for blk, cps, u in ((256, 8, 2), (256, 8, 4), (512, 4, 2), (512, 4, 4), (256, 6, 4), (256, 12, 2), (512, 2, 4), (1024, 2, 2), (256, 16, 1)):
        ctx.set_tuning("Stream_DOT", blk, cps, u)
        report(f"dot 2^28 block {blk} x {cps} CTAs/SM, {u} vectors", 16 * n, time_ms(lambda: ctx.stream_dot(a, b, o), 20))
    ctx.reset_tuning("Stream_DOT")
    n = 1 << 27
    for blk, cps, u in ((256, 4, 8), (256, 8, 4), (512, 4, 4), (512, 2, 8), (256, 8, 8), (512, 4, 8), (1024, 2, 4), (256, 6, 8)):
        ctx.set_tuning("Algorithm_REDUCE_SUM", blk, cps, u)
        report(f"reduce_sum 2^27 block {blk} x {cps} CTAs/SM, {u} vectors", 8 * n, time_ms(lambda: ctx.reduce_sum(a, o, n=n), 50))
    ctx.reset_tuning("Algorithm_REDUCE_SUM")
    del a, b

if out_path:
    json.dump(res, open(out_path, "w"), indent=1)

#!/usr/bin/env python
"""Small invocations of every kernel that hand-rolls synchronisation, for compute-sanitizer:

    compute-sanitizer --tool racecheck|synccheck|memcheck python tools/sanitize_targets.py [scan indexlist pa sort halo reduce ...]

scan / indexlist use the smallest n that still takes the TMA-staged, warp-specialised kernels (n >= 2 * 8192 * SMs) plus a
small n for the register-staged ones; pa = the three bulk-async ring kernels; sort = onesweep (general and uniform tiles) +
both histograms; halo = item-list kernels (one-launch pack+unpack, one-launch exchange) and the two-launch exchange with
4 ranks on one GPU (release/acquire flags); reduce = the ticketed grid fold.  Every result is also checked (exact data)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402

which = set(sys.argv[1:]) or {"scan", "indexlist", "pa", "sort", "halo", "reduce"}
ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")
sms = ctx.sm_count
ok = True


def check(name, cond):
    global ok
    ok = ok and bool(cond)
    print(f"{name}: {'ok' if cond else 'WRONG RESULT'}", flush=True)


if "reduce" in which:
    for n in (1000, 300001):
        a = torch.randint(-9, 9, (n,), device="cuda").to(torch.float64)
        o = torch.zeros(1, **f64)
        for _ in range(2):
            ctx.stream_dot(a, a, o); torch.cuda.synchronize()
            check(f"dot n={n}", o.item() == float((a * a).sum().item()))
            ctx.reduce_sum(a, o); torch.cuda.synchronize()
            check(f"reduce_sum n={n}", o.item() == float(a.sum().item()))

if "scan" in which:
    for n in (5000, 2 * 8192 * sms + 16 * 5 + 3):
        x = torch.randint(0, 5, (n,), device="cuda").to(torch.float64)
        y = torch.empty_like(x)
        for _ in range(2):
            ctx.scan_exclusive(x, y); torch.cuda.synchronize()
            check(f"scan n={n}", torch.equal(y, torch.cumsum(x, 0) - x))

if "indexlist" in which:
    for n in (5000, 2 * 8192 * sms + 16 * 5 + 3):
        x = torch.randint(-3, 4, (n,), device="cuda").to(torch.float64)
        lst = torch.full((n,), -1, dtype=torch.int32, device="cuda"); ln = torch.zeros(1, dtype=torch.int64, device="cuda")
        for _ in range(2):
            ctx.indexlist(x, lst, ln); torch.cuda.synchronize()
            want = torch.nonzero(x < 0).flatten().to(torch.int32)
            check(f"indexlist n={n}", ln.item() == want.numel() and torch.equal(lst[:want.numel()], want))

if "pa" in which:
    one = lambda m: torch.ones(m, **f64)
    for NE in (8, 1003):
        Y = torch.zeros(64 * NE, **f64)
        ctx.mass3dpa(one(20), one(20), one(125 * NE), one(64 * NE), Y, NE); torch.cuda.synchronize()
        check(f"mass3dpa NE={NE}", bool((Y == 8000.0).all()))
        Y = torch.zeros(27 * NE, **f64)
        ctx.convection3dpa(one(12), one(12), one(12), one(192 * NE), one(27 * NE), Y, NE); torch.cuda.synchronize()
        check(f"convection3dpa NE={NE}", bool((Y == 5184.0).all()))
        Y = torch.zeros(27 * NE, **f64)
        ctx.diffusion3dpa(one(12), one(12), one(384 * NE), one(27 * NE), Y, NE); torch.cuda.synchronize()
        check(f"diffusion3dpa NE={NE}", abs(float(Y.mean().item()) - 576.0) < 1e-9)

if "sort" in which:
    for n in (5 * 8192 + 77, 100003):
        for kind in ("random", "uniform_tiles"):
            x = torch.rand(n, **f64) if kind == "random" else torch.full((n,), 0.37, **f64)
            v = torch.arange(n, **f64)
            scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 32, **f64)
            for hist in (4, 8):
                ctx.set_tuning("Algorithm_SORT", -1, -1, hist); ctx.set_tuning("Algorithm_SORTPAIRS", -1, -1, hist)
                k = x.clone(); ctx.sort_keys(k, scratch); torch.cuda.synchronize()
                check(f"sort keys n={n} {kind} hist={hist}", torch.equal(k, torch.sort(x)[0]))
                k = x.clone(); w = v.clone(); ctx.sort_pairs(k, w, scratch); torch.cuda.synchronize()
                sk, si = torch.sort(x, stable=True)
                check(f"sort pairs n={n} {kind} hist={hist}", torch.equal(k, sk) and torch.equal(w, si.to(torch.float64)))
    ctx.reset_tuning("Algorithm_SORT"); ctx.reset_tuning("Algorithm_SORTPAIRS")

if "halo" in which:
    g, nv = 48, 3
    plan = ctx.halo_plan((g, g, g), 1, nv)
    mk = lambda: [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
    e = g + 2
    idx = torch.arange(e, device="cuda"); src = ((idx - 1) % g) + 1
    want = (src.view(e, 1, 1) * e * e + src.view(1, e, 1) * e + src.view(1, 1, e)).to(torch.float64)
    for unroll in (1, 2, 3):                                         # one launch (two phases) / pack launch + unpack launch / one launch (progressive)
        vars_ = mk()
        pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
        ub = [torch.full((nv * nb["unpack_len"],), 7.0, **f64) for nb in plan.neighbors]
        plan.bind(vars_, pb, ub)
        ctx.set_tuning("Comm_HALO_PACKING_FUSED", -1, -1, 1 if unroll == 3 else unroll)
        for _ in range(2):
            plan.pack_unpack()
        torch.cuda.synchronize()
        a = vars_[0].view(e, e, e)
        check(f"halo pack+unpack unroll={unroll}", bool((a[0] == 7.0).all()) and bool((a[1:-1, 1:-1, 1:-1] == mk()[0].view(e, e, e)[1:-1, 1:-1, 1:-1]).all())
              and torch.equal(pb[1][:plan.neighbors[1]["pack_len"]], mk()[0].view(e, e, e)[1:-1, 1:-1, g].flatten()))
        vars_ = mk()
        plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
        ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", -1, -1, unroll)
        for _ in range(3):
            plan.exchange()
        torch.cuda.synchronize(); plan.status()
        check(f"halo exchange 1 rank unroll={unroll}", all(torch.equal(vars_[v].view(e, e, e), want + v) for v in range(nv)))
    ctx.reset_tuning("Comm_HALO_PACKING_FUSED"); ctx.reset_tuning("Comm_HALO_EXCHANGE_FUSED")
    plan.close()
    # 4 ranks (2 x 2 x 1) on this one GPU, windows connected by plain pointers: pack + signal on every rank, then wait + unpack
    pd, P, g = (2, 2, 1), 4, 24
    plans, dv, wins = [], [], []
    for r in range(P):
        pl = ctx.halo_plan((g, g, g), 1, nv, r, pd)
        vs = [torch.arange(pl.var_size, **f64) + v for v in range(nv)]
        w, _, _ = pl.window(vs, want_handle=False)
        plans.append(pl); dv.append(vs); wins.append(w)
    for pl in plans:
        pl.connect_ptrs(wins)
    for _ in range(3):
        for pl in plans: pl.exchange_pack()
        for pl in plans: pl.exchange_unpack()
    torch.cuda.synchronize()
    e = g + 2
    idx = torch.arange(e, device="cuda"); src = ((idx - 1) % g) + 1
    want = (src.view(e, 1, 1) * e * e + src.view(1, e, 1) * e + src.view(1, 1, e)).to(torch.float64)
    good = True
    for r in range(P):
        plans[r].status()
        good = good and all(torch.equal(dv[r][v].view(e, e, e), want + v) for v in range(nv))
    check("halo exchange 4 ranks on one GPU (two launches per rank)", good)
    for pl in plans: pl.close()

print("SANITIZE_TARGETS", "PASS" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)

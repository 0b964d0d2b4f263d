#!/usr/bin/env python
"""Quick A/B timings on one B200 (run under gpurun): python tools/time_quick.py scan mass halo ..."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402

which = set(sys.argv[1:]) or {"scan", "pa", "halo", "sort"}
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


def report(name, bytes_, ms, **kw):
    res[name] = dict(ms=ms, gbs=bytes_ / ms / 1e6, **kw)
    print(f"{name:48s} {ms:9.4f} ms {bytes_ / ms / 1e6:9.1f} GB/s {kw if kw else ''}", flush=True)


if "scan" in which:
    n = 1 << 27
    x = torch.rand(n, **f64); y = torch.empty(n, **f64)
    for tune in [(512, 1, 4), (512, 2, 4), (512, 4, 4), (256, 2, 4), (256, 4, 4), (256, 8, 4), (256, 4, 2), (512, 2, 2), (1024, 1, 2), (128, 4, 4)]:
        try:
            ctx.set_tuning("Algorithm_SCAN", *tune)
            report(f"scan{tune}", 16 * n, time_ms(lambda: ctx.scan_exclusive(x, y)))
        except Exception as e:
            print("scan", tune, "failed", e)
    del x, y

if "pa" in which:
    NE = 4000000
    one = lambda m: torch.ones(m, **f64)
    B, Bt, D, X, Y = one(20), one(20), one(125 * NE), one(64 * NE), one(64 * NE)
    for var, shape in ((1, "8/32/12 one D stage (default)"), (30, "8/32/8 two D stages"), (27, "8/32/11 one D stage"), (28, "8/32/13 one D stage"), (29, "16/64/6 one D stage")):
        ctx.set_tuning("Apps_MASS3DPA", -1, -1, var)
        report(f"mass3dpa E/BLOCK/CTAs={shape}", 2536 * NE, time_ms(lambda: ctx.mass3dpa(B, Bt, D, X, Y, NE), 10))
    ctx.set_tuning("Apps_MASS3DPA", -1, -1, 1)
    del D, X, Y
    B, G, D, X, Y = one(12), one(12), one(192 * NE), one(27 * NE), one(27 * NE)
    for var, shape in ((1, "8/128/2/5 (default)"),):
        ctx.set_tuning("Apps_CONVECTION3DPA", -1, -1, var)
        report(f"convection3dpa E/BLOCK/S/CTAs={shape}", 2184 * NE, time_ms(lambda: ctx.convection3dpa(B, B, G, D, X, Y, NE), 10))
    ctx.set_tuning("Apps_CONVECTION3DPA", -1, -1, 1)
    del D, X, Y

if "ltimes" in which:
    nz = 500000
    phi = torch.zeros(800 * nz, **f64); psi = torch.rand(2048 * nz, **f64); ell = torch.rand(1600, **f64)
    for var, label in ((4, "register prefetch (default)"), (8, "psi ring 3 stages, row copies"), (5, "ring 3 stages, one 4 KB copy"), (6, "ring 4 stages one copy"), (7, "ring 6 stages one copy")):
        for cps in (2,):
            ctx.set_tuning("Apps_LTIMES", -1, cps, var)
            ms = time_ms(lambda: ctx.ltimes(phi, ell, psi, 64, 32, 25, nz), 10)
            report(f"ltimes {label} cps={cps}", 912 * 32 * nz, ms, tflops=3200 * 32 * nz / ms / 1e9)
    ctx.set_tuning("Apps_LTIMES", -1, 2, 4)
    del phi, psi

if "sort_hist" in which:
    # SORT / SORTPAIRS with the lane-private histogram kernel (the default since round 2) against the shared-bin histogram
    # (tuning unroll 8): parity of the sorted output, then A/B/A/B.  First run: profiles/r02_a_optin.log.
    n = 1 << 27
    x = torch.randint(0, 2**31 - 1, (n,), device="cuda").to(torch.float64).div_(2147483647.0)      # rand()/RAND_MAX
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 32, **f64)
    for m_ in (n, 1000003, 4097):
        outs = []
        for var in (8, 4):
            ctx.set_tuning("Algorithm_SORT", -1, -1, var); ctx.set_tuning("Algorithm_SORTPAIRS", -1, -1, var)
            k = x[:m_].clone(); ctx.sort_keys(k, scratch, n=m_)
            kb, vb = torch.empty(m_ + 1, **f64), torch.empty(m_ + 1, **f64)
            k2, v2 = (kb[1:], vb[1:]) if m_ < n else (kb[:m_], vb[:m_])                # [1:]: key pointer not 32-byte aligned
            k2.copy_(x[:m_]); v2.copy_(x[:m_])
            ctx.sort_pairs(k2, v2, scratch, n=m_)
            outs.append((k, k2, v2))
        ok = all(bool(torch.equal(a, b)) for a, b in zip(outs[0], outs[1])) and bool((outs[0][0][1:] >= outs[0][0][:-1]).all())
        print(f"sort lane-private histogram parity n={m_}: {'OK' if ok else 'FAILED'}", flush=True)
        res[f"sort_hist parity {m_}"] = ok
        del outs
    y = torch.empty_like(x)
    for rnd in range(2):
        for var, label in ((8, "shared-bin histogram"), (4, "lane-private histogram (default)")):
            ctx.set_tuning("Algorithm_SORT", -1, -1, var)
            ms = time_ms(lambda: ctx.sort_keys(y, scratch), 5, 2, setup=lambda: y.copy_(x))
            report(f"sort keys, {label} round {rnd}", 16 * n, ms, mkeys_per_s=n / ms / 1e3)
    ctx.reset_tuning("Algorithm_SORT"); ctx.reset_tuning("Algorithm_SORTPAIRS")
    del x, y, scratch

if "ltimes_line" in which:
    # A fragments owned line-major (the default) against the row-chunk mapping (unroll 10): parity, then A/B/A/B
    nz0 = 37
    g = torch.Generator(device="cuda").manual_seed(9)
    phi0 = torch.randint(-5, 6, (nz0 * 32, 25), generator=g, **{**f64, "dtype": torch.int64}).to(torch.float64)
    ell0 = torch.randint(-3, 4, (25, 64), generator=g, device="cuda").to(torch.float64)
    psi0 = torch.randint(-3, 4, (nz0 * 32, 64), generator=g, device="cuda").to(torch.float64)
    want = phi0 + psi0 @ ell0.t()
    for var in (10, 4):
        got = phi0.clone().reshape(-1)
        ctx.set_tuning("Apps_LTIMES", -1, 2, var)
        ctx.ltimes(got, ell0.reshape(-1).contiguous(), psi0.reshape(-1).contiguous(), 64, 32, 25, nz0)
        ok = bool(torch.equal(got.reshape(-1, 25), want))
        print(f"ltimes variant {var}: integer-valued parity {'OK' if ok else 'FAILED'}", flush=True)
        res[f"ltimes parity variant {var}"] = ok
    nz = 500000
    phi = torch.zeros(800 * nz, **f64); psi = torch.rand(2048 * nz, **f64); ell = torch.rand(1600, **f64)
    for rnd in range(2):
        for var, label in ((10, "row chunks"), (4, "line-major fragments (default)")):
            ctx.set_tuning("Apps_LTIMES", -1, 2, var)
            ms = time_ms(lambda: ctx.ltimes(phi, ell, psi, 64, 32, 25, nz), 10)
            report(f"ltimes {label} round {rnd}", 912 * 32 * nz, ms, tflops=3200 * 32 * nz / ms / 1e9)
    ctx.reset_tuning("Apps_LTIMES")
    del phi, psi

if "mass_line" in which:
    # line-major X / Y accesses (the default = unroll 32: 11 CTAs per SM; 31: 12) against slab-per-thread loads (unroll 24):
    # parity against the slab-per-thread kernel on integer-valued data, then A/B/A/B
    for NE0 in (1024, 1029, 40000):
        g = torch.Generator(device="cuda").manual_seed(NE0)
        ri = lambda n, lo, hi: torch.randint(lo, hi, (n,), generator=g, device="cuda").to(torch.float64)
        B0, Bt0, D0, X0, Y0 = ri(20, -2, 3), ri(20, -2, 3), ri(125 * NE0, -3, 4), ri(64 * NE0, -3, 4), ri(64 * NE0, -3, 4)
        outs = {}
        for var in (24, 1, 31):
            ctx.set_tuning("Apps_MASS3DPA", -1, -1, var)
            Yv = Y0.clone()
            ctx.mass3dpa(B0, Bt0, D0, X0, Yv, NE0); ctx.mass3dpa(B0, Bt0, D0, X0, Yv, NE0)
            outs[var] = Yv
        ok = bool(torch.equal(outs[24], outs[1]) and torch.equal(outs[24], outs[31]) and not torch.equal(outs[24], Y0))
        print(f"mass3dpa line-major parity NE={NE0}: {'OK' if ok else 'FAILED'}", flush=True)
        res[f"mass parity NE={NE0}"] = ok
    NE = 4000000
    one = lambda m: torch.ones(m, **f64)
    B, Bt, D, X, Y = one(20), one(20), one(125 * NE), one(64 * NE), torch.zeros(64 * NE, **f64)
    for rnd in range(2):
        for var, label in ((24, "slab per thread, 12 CTAs/SM"), (31, "line-major, 12 CTAs/SM"), (1, "line-major, 11 CTAs/SM (default)"), (33, "line-major, 10 CTAs/SM"), (34, "line-major, 9 CTAs/SM")):
            ctx.set_tuning("Apps_MASS3DPA", -1, -1, var)
            ms = time_ms(lambda: ctx.mass3dpa(B, Bt, D, X, Y, NE), 10)
            report(f"mass3dpa {label} round {rnd}", 2536 * NE, ms)
    ctx.reset_tuning("Apps_MASS3DPA")
    del D, X, Y

if "halo" in which:
    for g in (512,):
        nv = 3
        plan = ctx.halo_plan((g, g, g), 1, nv)
        vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
        pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
        ub = [torch.zeros(nv * nb["unpack_len"], **f64) for nb in plan.neighbors]
        plan.bind(vars_, pb, ub)
        ne = sum(nb["pack_len"] for nb in plan.neighbors) * nv
        plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
        def graph_ms(body, reps=100):
            body(); torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                for _ in range(reps):
                    body()
            g_.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        # `unroll`: 1 = no L2 hints, 2 = exchange as two launches (no hints), 4 = hints; `block_size` 128 = round-robin chunks
        for hint, blk in ((1, 256), (4, 256), (1, 128), (4, 128)):
          for cps in (4, 8):
            ctx.set_tuning("Comm_HALO_PACKING_FUSED", blk, cps, hint)
            tag = f"cps={cps} hint={int(hint == 4)} rr={int(blk == 128)}"
            report(f"halo{g} pack+unpack [graph] {tag}", 40 * ne, graph_ms(lambda: (plan.pack(), plan.unpack())))
            for xu in ((1, 2) if hint == 1 else (4,)):      # exchange: 1 = one fused launch, 2 / 4 = two launches without / with hints
                ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", blk, cps, xu)
                report(f"halo{g} exchange [graph] {tag} launches={1 if xu == 1 else 2}", 56 * ne, graph_ms(plan.exchange))
        plan.status()
        plan.close()
        del vars_, pb, ub

if "sort" in which:
    n = 1 << 27
    src = torch.randint(0, 2**31 - 1, (n,), device="cuda").to(torch.float64).div_(2147483647.0)
    k = torch.empty_like(src); v = torch.empty_like(src)
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 32, **f64)
    ms = time_ms(lambda: ctx.sort_keys(k, scratch), 5, 2, setup=lambda: k.copy_(src))
    report("sort_keys", 16 * n, ms, mkeys_s=n / ms / 1e3)
    ms = time_ms(lambda: ctx.sort_pairs(k, v, scratch), 5, 2, setup=lambda: (k.copy_(src), v.copy_(src)))
    report("sort_pairs", 32 * n, ms, mkeys_s=n / ms / 1e3)
    ms = time_ms(lambda: torch.sort(src), 5, 2)
    report("torch.sort (CUB pairs: keys + int64 indices)", 32 * n, ms, mkeys_s=n / ms / 1e3)

if "indexlist" in which:
    n = 1 << 27
    x = torch.randn(n, **f64); lst = torch.empty(n, dtype=torch.int32, device="cuda")
    ln = torch.zeros(1, dtype=torch.int64, device="cuda")
    for cps in (2, 3, 4):
        ctx.set_tuning("Basic_INDEXLIST", -1, cps, -1)
        ms = time_ms(lambda: ctx.indexlist(x, lst, ln))
        report(f"indexlist cps={cps}", 8 * n + 4 * int(ln.item()), ms)
    del x, lst

if "gemm" in which:
    shapes = [("64x64x16 s3 (auto small)", 64, 4), ("128x64x16 s3 (auto large)", 96, 4), ("64x64x16 s2 5cta", 64, 20), ("64x64x16 s2 6cta", 64, 21),
              ("64x64x32 s2 3cta", 64, 22), ("128x64x16 s2 3cta", 64, 23), ("64x64x16 s3 4cta", 64, 24), ("64x64x16 s3 8 warps of 32x16", 64, 25)]
    for ni in (1000, 4096, 8192):
        nj, nk = ni, int(1.2 * ni)
        A = torch.rand(ni * nk, **f64); B = torch.rand(nk * nj, **f64); C = torch.empty(ni * nj, **f64)
        for label, t, u in shapes:
            ctx.set_tuning("Polybench_GEMM", t, -1, u)
            ms = time_ms(lambda: ctx.polybench_gemm(A, B, C, ni, nj, nk, 0.62), 10 if ni < 8192 else 4)
            report(f"gemm {ni}x{nj}x{nk} {label}", 8 * (ni * nk + nk * nj + ni * nj), ms, tflops=2.0 * ni * nj * nk / ms / 1e9)
        ms = time_ms(lambda: torch.mm(A.view(ni, nk), B.view(nk, nj), out=C.view(ni, nj)), 10 if ni < 8192 else 4)
        report(f"gemm {ni} cuBLAS dgemm (incumbent)", 8 * (ni * nk + nk * nj + ni * nj), ms, tflops=2.0 * ni * nj * nk / ms / 1e9)
        ctx.set_tuning("Polybench_GEMM", 256, -1, 4)
        del A, B, C

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/time_quick.json", "w"), indent=1)

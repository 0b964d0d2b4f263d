#!/usr/bin/env python
"""Halo timing vs chunk dealing across fresh allocations (address placement) in ONE process.  Run under gpurun."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context
ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")

def graph_ms(body, reps=100):
    body(); torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(reps): body()
    g_.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3

keep = []
for trial in range(2):
    if trial: keep.append(torch.empty((trial * 37 + 5) << 20, dtype=torch.uint8, device="cuda"))   # shift the addresses
    g, nv = 512, 3
    plan = ctx.halo_plan((g, g, g), 1, nv)
    vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
    pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
    ub = [torch.zeros(nv * nb["unpack_len"], **f64) for nb in plan.neighbors]
    plan.bind(vars_, pb, ub)
    plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
    out = []
    for blk, cps in ((192, 4), (160, 4), (160, 8), (192, 2), (192, 3)):
        ctx.set_tuning("Comm_HALO_PACKING_FUSED", blk, cps, 1)
        out.append(f"pk {'rr' if blk == 128 else ('rv' if blk == 192 else ('rrv' if blk == 160 else 'ct'))}{cps} {graph_ms(lambda: (plan.pack(), plan.unpack())):6.1f}")
    for blk, cps, xu in ((192, 4, 2), (160, 4, 2), (160, 8, 2), (192, 3, 2)):
        ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", blk, cps, xu)
        out.append(f"ex {'rr' if blk == 128 else ('rv' if blk == 192 else ('rrv' if blk == 160 else 'ct'))}{cps}/{'1L' if xu == 1 else '2L'} {graph_ms(plan.exchange):6.1f}")
    print(f"trial {trial} var0@{vars_[0].data_ptr():#x}: " + " | ".join(out), flush=True)
    plan.status(); plan.close()
    del vars_, pb, ub, plan
    torch.cuda.empty_cache()

#!/usr/bin/env python
"""Which faces cost what: pack/unpack time per face orientation at 512^3 (run under gpurun)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context

ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")
g, nv = int(os.environ.get("G", 512)), 3
plan = ctx.halo_plan((g, g, g), 1, nv)
vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]

def time_ms(fn, reps=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, ls in (("x faces", (0, 1)), ("y faces", (2, 3)), ("z faces", (4, 5)), ("edges+corners", tuple(range(6, 26))), ("all", tuple(range(26)))):
    bufs = {l: torch.zeros(nv * plan.neighbors[l]["pack_len"], **f64) for l in ls}
    psegs, usegs = [], []
    for l in ls:
        nb = plan.neighbors[l]
        for v in range(nv):
            psegs.append((bufs[l].data_ptr() + 8 * v * nb["pack_len"], nb["d_pack_list"], vars_[v], nb["pack_len"], l))
            usegs.append((bufs[l].data_ptr() + 8 * v * nb["unpack_len"], nb["d_unpack_list"], vars_[v], nb["unpack_len"], l))
    pw, uw = ctx.halo_worklist(psegs), ctx.halo_worklist(usegs)
    ne = sum(plan.neighbors[l]["pack_len"] for l in ls) * nv
    tp, tu = time_ms(lambda: ctx.halo_pack(pw)), time_ms(lambda: ctx.halo_unpack(uw))
    print(f"{name:14s} elems {ne:9d}  pack {tp*1e3:7.1f} us {20*ne/tp/1e6:7.0f} GB/s   unpack {tu*1e3:7.1f} us {20*ne/tu/1e6:7.0f} GB/s", flush=True)

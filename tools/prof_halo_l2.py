#!/usr/bin/env python
"""Warm-cache counters of the halo kernels per face orientation, with and without the L2 eviction-priority
hints (run under `ncu --cache-control none --clock-control none --metrics ...`, see tools/gpu_exp2.sh)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context

ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")
g, nv = int(os.environ.get("G", 512)), 3
plan = ctx.halo_plan((g, g, g), 1, nv)
vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
for name, ls in (("x", (0, 1)), ("y", (2, 3)), ("all", tuple(range(26)))):
    bufs = {l: torch.zeros(nv * plan.neighbors[l]["pack_len"], **f64) for l in ls}
    psegs, usegs = [], []
    for l in ls:
        nb = plan.neighbors[l]
        for v in range(nv):
            psegs.append((bufs[l].data_ptr() + 8 * v * nb["pack_len"], nb["d_pack_list"], vars_[v], nb["pack_len"], l))
            usegs.append((bufs[l].data_ptr() + 8 * v * nb["unpack_len"], nb["d_unpack_list"], vars_[v], nb["unpack_len"], l))
    pw, uw = ctx.halo_worklist(psegs), ctx.halo_worklist(usegs)
    for hint in (1, 4):
        ctx.set_tuning("Comm_HALO_PACKING_FUSED", -1, 4, hint)
        for _ in range(3): ctx.halo_pack(pw)
        for _ in range(3): ctx.halo_unpack(uw)
        for _ in range(2): ctx.halo_pack(pw); ctx.halo_unpack(uw)
        torch.cuda.synchronize()
        print("done", name, hint, flush=True)

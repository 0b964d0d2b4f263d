#!/usr/bin/env python
"""DRAM traffic of pack / unpack launches with the packs walking forward (block_size 256) or backwards (192), warm L2
(run under `ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context

ctx = Context(0)
f64 = dict(dtype=torch.float64, device="cuda")
g, nv = 512, 3
plan = ctx.halo_plan((g, g, g), 1, nv)
vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
ub = [torch.zeros(nv * nb["unpack_len"], **f64) for nb in plan.neighbors]
plan.bind(vars_, pb, ub)
for blk in (256, 192):
    ctx.set_tuning("Comm_HALO_PACKING_FUSED", blk, 4, 1)
    for _ in range(4):
        plan.pack(); plan.unpack()
    torch.cuda.synchronize()
    print("done", blk, flush=True)

#!/usr/bin/env python
"""The halo launches of HEAD for ncu (warm L2, like the rep loop):
    ncu --set full --cache-control none --clock-control none --import-source on -k regex:'halo_items_kernel|halo_kernel' -c 17 -o out python tools/prof_halo_r02.py
3 reps of each one-launch order, 2 of the two-launch form, at 512^3 x 3 variables; then 4 of the one-launch exchange (1 rank)."""
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
K = "Comm_HALO_PACKING_FUSED"
for order in (1, 3, 5):                 # x units mixed in / x units first / two phases: 3 launches each (the last is warm)
    ctx.set_tuning(K, 192, 4, order)
    for _ in range(3):
        plan.pack_unpack()
    torch.cuda.synchronize()
ctx.set_tuning(K, 192, 4, 2)
for _ in range(2):
    plan.pack(); plan.unpack()
torch.cuda.synchronize()
plan.window(vars_, want_handle=False); plan.connect_ptrs([0])
ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", 192, 4, 1)
for _ in range(4):
    plan.exchange()
torch.cuda.synchronize()
plan.status()
print("done", flush=True)

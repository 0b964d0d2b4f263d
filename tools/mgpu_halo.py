#!/usr/bin/env python
"""HALO_EXCHANGE_FUSED time per rep on the N-rank grid for several launch tunings (torchrun, one rank per GPU):
block_size 256/128 = contiguous / round-robin chunks, unroll 1 = one fused launch, 2 = pack launch + unpack launch."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context
from rajaperf_b200.dist import rank_grid

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = Context(local)
f64 = dict(dtype=torch.float64, device=dev)
g, nv = int(os.environ.get("G", 512)), 3
plan = ctx.halo_plan((g, g, g), 1, nv, rank, rank_grid(world))
vars_ = [torch.arange(plan.var_size, **f64) + v for v in range(nv)]
_, _, handle = plan.window(vars_)
if world > 1:
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    plan.connect(handles)
    dist.barrier()
else:
    plan.connect_ptrs([0])

def graph_ms(body, reps=100):
    body(); torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(reps): body()
    g_.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], **f64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        best = min(best, ms)
    return best

res = {}
for blk, cps, xu in ((256, 4, 2), (192, 4, 2), (192, 8, 2)):
    ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", blk, cps, xu)
    ms = graph_ms(plan.exchange)
    plan.status()
    res[f"{'rr' if blk == 128 else ('rv' if blk == 192 else 'ct')}{cps}/{'1L' if xu == 1 else '2L'}"] = round(ms * 1e3, 1)
if rank == 0:
    print(json.dumps({"n_gpus": world, "rank_grid": rank_grid(world), "cells_per_gpu": g, "us_per_rep": res}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

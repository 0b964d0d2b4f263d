#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
  * HALO_EXCHANGE_FUSED on the default rank grid vs the CPU simulation of the same grid (bit-exact), in both launch forms
    (pack launch + unpack launch; ONE launch in two phases; ONE launch with progressive signalling)
  * global DOT / REDUCE_SUM: shards + one all-reduced scalar vs the oracle on the whole array.
Rank 0 prints `MGPU_CHECK PASS|FAIL ...` and one JSON line with what was compared (tools/gpu_call.sh mgpu_check:P keeps both)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                      # noqa: E402
import suite_data as sd            # noqa: E402
from rajaperf_b200 import Context  # noqa: E402
from rajaperf_b200 import dist as rdist  # noqa: E402
from test_comm_gpu import simulate_exchange  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
pd = rdist.rank_grid(world)
ok = True
cases = []
for unroll, dims, hw, nv, reps in [(u,) + c for u in (2, 1, 3) for c in (((6, 6, 6), 1, 3, 3), ((40, 40, 40), 2, 2, 4), ((128, 128, 128), 1, 3, 5),
                                                                        ((252, 252, 252), 1, 3, 2))]:
    ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", unroll=unroll)
    plan = ctx.halo_plan(dims, hw, nv, rank, pd)
    vs = [torch.arange(plan.var_size, dtype=torch.float64, device="cuda") + v for v in range(nv)]
    rdist.connect_halo_plan(plan, vs)
    for _ in range(reps):
        plan.exchange()
    torch.cuda.synchronize()
    plan.status()
    dist.barrier()
    ref = simulate_exchange(dims, hw, nv, pd, reps)[rank]
    for v in range(nv):
        same = np.array_equal(vs[v].cpu().numpy().view(np.int64), ref[v].view(np.int64))
        ok &= same
        if not same:
            print(f"rank {rank}: halo mismatch dims={dims} var={v}", flush=True)
    dist.barrier()
    plan.close()
    cases.append({"form": {2: "pack launch + unpack launch", 1: "one launch, two phases", 3: "one launch, progressive"}[unroll], "cells": list(dims), "halo_width": hw, "vars": nv, "reps": reps})
ctx.reset_tuning("Comm_HALO_EXCHANGE_FUSED")

n = 3000001
d = sd.stream_dot(n)
b, e = rdist.shard_range(n, rank, world)
a_s, b_s = torch.from_numpy(d["a"][b:e].copy()).cuda(), torch.from_numpy(d["b"][b:e].copy()).cuda()
out = torch.zeros(1, dtype=torch.float64, device="cuda")
ctx.stream_dot(a_s, b_s, out)
rdist.allreduce_scalar(out)
ref = oracle.lib().orc_stream_dot(d["a"], d["b"], n, 0.0)
if abs(out.item() - ref) > 1e-7:
    ok = False
    print(f"rank {rank}: dot {out.item()} vs {ref}", flush=True)
ctx.reduce_sum(a_s, out)
rdist.allreduce_scalar(out)
ref = oracle.lib().orc_reduce_sum(d["a"], n, 0.0)
if abs(out.item() - ref) > 1e-7 * max(1.0, abs(ref)) * 1e-2:
    ok = False
    print(f"rank {rank}: sum {out.item()} vs {ref}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU_CHECK", "PASS" if flag.item() == 1 else "FAIL", f"world={world} grid={pd}", flush=True)
    print(json.dumps({"mgpu_check": "PASS" if flag.item() == 1 else "FAIL", "world": world, "rank_grid": pd,
                      "halo_exchange_cases_bit_exact_vs_cpu_simulation": cases,
                      "sharded_dot_and_reduce_sum_vs_oracle": {"n": n, "tolerance_abs": 1e-7}}), flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)

#!/usr/bin/env python
"""Run ONE kernel a few times for ncu (python tools/prof_one.py <kernel> [n])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 27)
reps = int(os.environ.get("REPS", 3))
ctx = Context(0)
x = torch.rand(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
if which == "scan":
    for _ in range(reps):
        ctx.scan_exclusive(x, y)
elif which == "sort":
    scratch = torch.empty(ctx.sort_scratch_bytes(n, False) // 8 + 1, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        y.copy_(x)
        ctx.sort_keys(y, scratch)
elif which == "sortpairs":
    v = x.clone()
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 1, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        y.copy_(x)
        ctx.sort_pairs(y, v, scratch)
elif which == "triad":
    z = torch.rand(n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        ctx.stream_triad(y, x, z, 0.3)
elif which == "dot":
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        ctx.stream_dot(y, x, out)
torch.cuda.synchronize()

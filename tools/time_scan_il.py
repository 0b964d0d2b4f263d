#!/usr/bin/env python
"""SCAN and INDEXLIST at 2^27 (TMA paths); descriptor spacing comes from RPB200_SCAN_DSTRIDE / RPB200_IL_DSTRIDE."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context

ctx = Context(0)
n = 1 << 27
f64 = dict(dtype=torch.float64, device="cuda")

def time_ms(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

x = torch.rand(n, **f64); y = torch.empty(n, **f64)
ms = time_ms(lambda: ctx.scan_exclusive(x, y))
print(f"scan      dstride={os.environ.get('RPB200_SCAN_DSTRIDE', 'default')}: {ms:.4f} ms {16 * n / ms / 1e6:.0f} GB/s", flush=True)
x = torch.randn(n, **f64); lst = torch.empty(n, dtype=torch.int32, device="cuda"); ln = torch.zeros(1, dtype=torch.int64, device="cuda")
ms = time_ms(lambda: ctx.indexlist(x, lst, ln))
print(f"indexlist dstride={os.environ.get('RPB200_IL_DSTRIDE', 'default')}: {ms:.4f} ms {(8 * n + 4 * int(ln.item())) / ms / 1e6:.0f} GB/s", flush=True)

#!/usr/bin/env python
"""Quick timing of SCAN / SORT / SORTPAIRS at BASELINE sizes vs torch (CUB-backed) -- run under gpurun."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402


def time_ms(fn, reps=10, warm=3, setup=None):
    for _ in range(warm):
        if setup: setup()
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if setup: setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    n = int(os.environ.get("N", 1 << 27))
    ctx = Context(0)
    res = {"n": n}
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    for tune in [(256, 4, 4), (256, 8, 4), (512, 2, 4), (512, 4, 4), (256, 8, 2), (128, 8, 4), (128, 16, 4), (512, 4, 2), (256, 6, 4)]:
        ctx.set_tuning("Algorithm_SCAN", *tune)
        ms = time_ms(lambda: ctx.scan_exclusive(x, y))
        res[f"scan{tune}"] = {"ms": ms, "gbs": 16 * n / ms / 1e6}
        print("scan", tune, res[f"scan{tune}"], flush=True)
    ms = time_ms(lambda: torch.cumsum(x, 0, out=y))
    res["torch.cumsum"] = {"ms": ms, "gbs": 16 * n / ms / 1e6}
    print("torch.cumsum", res["torch.cumsum"], flush=True)

    src = torch.rand(n, dtype=torch.float64, device="cuda")
    k = torch.empty_like(src)
    v = torch.empty_like(src)
    scratch = torch.empty(ctx.sort_scratch_bytes(n, True) // 8 + 1, dtype=torch.float64, device="cuda")
    ms = time_ms(lambda: ctx.sort_keys(k, scratch), setup=lambda: k.copy_(src), reps=5)
    res["sort_keys"] = {"ms": ms, "gbs_nominal": 16 * n / ms / 1e6, "mkeys_s": n / ms / 1e3}
    print("sort_keys", res["sort_keys"], flush=True)
    ms = time_ms(lambda: ctx.sort_pairs(k, v, scratch), setup=lambda: (k.copy_(src), v.copy_(src)), reps=5)
    res["sort_pairs"] = {"ms": ms, "gbs_nominal": 32 * n / ms / 1e6, "mkeys_s": n / ms / 1e3}
    print("sort_pairs", res["sort_pairs"], flush=True)
    ms = time_ms(lambda: torch.sort(src), reps=5)
    res["torch.sort(keys+indices)"] = {"ms": ms, "mkeys_s": n / ms / 1e3}
    print("torch.sort", res["torch.sort(keys+indices)"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/time_algorithm.json", "w"), indent=1)


if __name__ == "__main__":
    main()

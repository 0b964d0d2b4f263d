#!/usr/bin/env python
"""Tuning sweep for the Stream group + REDUCE_SUM on one B200 (run under gpurun).
Writes gpurun_out/sweep_stream.json; the winners become the defaults in csrc/ctx.cu."""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402


def time_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(os.environ.get("SWEEP_N", 1 << 28))
    ctx = Context(0)
    a = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(0.1, 0.2)
    b = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(0.1, 0.2)
    c = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(0.1, 0.2)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    kernels = {
        "Stream_COPY": (16, lambda: ctx.stream_copy(c, a)),
        "Stream_MUL": (16, lambda: ctx.stream_mul(b, c, 0.3)),
        "Stream_ADD": (24, lambda: ctx.stream_add(c, a, b)),
        "Stream_TRIAD": (24, lambda: ctx.stream_triad(a, b, c, 0.3)),
        "Stream_DOT": (16, lambda: ctx.stream_dot(a, b, out)),
        "Algorithm_REDUCE_SUM": (8, lambda: ctx.reduce_sum(a, out)),
    }
    res = {"n": n, "torch": {}, "sweep": {}}
    res["torch"]["copy_"] = 16 * n / time_ms(lambda: c.copy_(a)) / 1e6
    res["torch"]["add"] = 24 * n / time_ms(lambda: torch.add(a, b, out=c)) / 1e6
    res["torch"]["dot"] = 16 * n / time_ms(lambda: torch.dot(a, b)) / 1e6
    res["torch"]["sum"] = 8 * n / time_ms(lambda: a.sum()) / 1e6
    print("torch GB/s", res["torch"], flush=True)
    for name, (bpe, fn) in kernels.items():
        rows = []
        reduce_like = name in ("Stream_DOT", "Algorithm_REDUCE_SUM")
        cps_opts = [1, 2, 4, 8, 16] if reduce_like else [0, 2, 4, 8, 16]
        for bs, cps, u in itertools.product([128, 256, 512], cps_opts, [1, 2, 4, 8]):
            if cps * bs > 2048:
                continue
            ctx.set_tuning(name, bs, cps, u)
            ms = time_ms(fn, reps=6, warm=2)
            rows.append({"block": bs, "ctas_per_sm": cps, "unroll": u, "ms": ms, "gbs": bpe * n / ms / 1e6})
        rows.sort(key=lambda r: -r["gbs"])
        res["sweep"][name] = rows
        print(name, "best:", rows[:4], "worst:", rows[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/sweep_stream.json", "w"), indent=1)


if __name__ == "__main__":
    main()

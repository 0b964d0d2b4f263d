#!/usr/bin/env python
"""Quick timing of the Apps kernels at BASELINE config #4 sizes -- run under gpurun."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajaperf_b200 import Context  # noqa: E402


def time_ms(fn, reps=8, warm=3):
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
    ctx = Context(0)
    NE = int(os.environ.get("NE", 4000000))
    one = lambda n: torch.ones(n, dtype=torch.float64, device="cuda")
    res = {}

    def report(name, bytes_, flops, ms):
        res[name] = {"ms": ms, "gbs": bytes_ / ms / 1e6, "gflops": flops / ms / 1e6}
        print(name, res[name], flush=True)

    B, Bt, D, X, Y = one(20), one(20), one(125 * NE), one(64 * NE), one(64 * NE)
    for cps in (0, 2, 4, 8):
        ctx.set_tuning("Apps_MASS3DPA", -1, cps, -1)
        report(f"mass3dpa cps={cps}", 2536 * NE, 5069 * NE, time_ms(lambda: ctx.mass3dpa(B, Bt, D, X, Y, NE)))
    del D, X, Y
    B, G, D, X, Y = one(12), one(12), one(384 * NE), one(27 * NE), one(27 * NE)
    for cps in (0, 2, 4, 8):
        ctx.set_tuning("Apps_DIFFUSION3DPA", -1, cps, -1)
        report(f"diffusion3dpa cps={cps}", 3720 * NE, 7065 * NE, time_ms(lambda: ctx.diffusion3dpa(B, G, D, X, Y, NE)))
    del D
    D = one(192 * NE)
    for cps in (0, 2, 4, 8):
        ctx.set_tuning("Apps_CONVECTION3DPA", -1, cps, -1)
        report(f"convection3dpa cps={cps}", 2184 * NE, 3683 * NE, time_ms(lambda: ctx.convection3dpa(B, Bt, G, D, X, Y, NE)))
    del D, X, Y
    nz = int(os.environ.get("NZ", 500000))
    phi = torch.zeros(800 * nz, dtype=torch.float64, device="cuda")
    psi = torch.rand(2048 * nz, dtype=torch.float64, device="cuda")
    ell = torch.rand(1600, dtype=torch.float64, device="cuda")
    for cps in (1, 2, 3, 4):
        ctx.set_tuning("Apps_LTIMES", -1, cps, -1)
        report(f"ltimes cps={cps}", 912 * 32 * nz, 3200 * 32 * nz, time_ms(lambda: ctx.ltimes(phi, ell, psi, 64, 32, 25, nz)))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/time_apps.json", "w"), indent=1)


if __name__ == "__main__":
    main()

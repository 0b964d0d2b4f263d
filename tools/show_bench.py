#!/usr/bin/env python
"""Pretty-print a bench.py JSON line: python tools/show_bench.py gpurun_out/xxx.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(f"value {d['value']:.1f} {d['unit']}  n_gpus {d['n_gpus']}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.1f}  clocks {d.get('clocks')}")
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"])
print("roofline", d.get("roofline"))
for k, v in (d.get("kernels") or {}).items():
    extra = f" {v['mkeys_per_s']:.0f} Mkeys/s" if "mkeys_per_s" in v else (f" {v['gflops']:.0f} GFLOP/s" if "gflops" in v else "")
    print(f"{k:28s} {v['ms']:9.4f} ms {v['gbs']:8.1f} GB/s  {v['frac_of_measured_peak']:.3f} of measured peak{extra}")
print("halo_exchange", {k: v for k, v in (d.get("halo_exchange") or {}).items() if k != "transport"})

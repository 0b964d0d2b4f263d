#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary committed under profiles/:
python tools/ncu_summary.py raw.csv out.md [traffic.json]"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default="-"):
    return r[col[name]] if name in col and r[col[name]] != "" else default


def to_bytes(v, unit):
    v = float(v)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)


short = lambda n: re.sub(r"\(.*", "", n.replace("void ", "").replace("<unnamed>::", ""))
out = ["| kernel | grid x block | regs | time (us, ncu cold) | DRAM read (MB) | DRAM write (MB) | DRAM GB/s | DRAM % of peak | warps active % | FP64 pipe % |",
       "|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for r in rows[2:]:
    name = short(g(r, "Kernel Name"))
    t_us = float(g(r, "gpu__time_duration.sum", 0)) * {"ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1)
    rd = to_bytes(g(r, "dram__bytes_read.sum", 0), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(g(r, "dram__bytes_write.sum", 0), units[col["dram__bytes_write.sum"]])
    out.append(f"| `{name}` | {g(r,'launch__grid_size')} x {g(r,'launch__block_size')} | {g(r,'launch__registers_per_thread')} | {t_us:.1f} | "
               f"{rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/t_us/1e3:.0f} | {float(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',0)):.1f} | "
               f"{float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active',0)):.1f} | {float(g(r,'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',0)):.1f} |")
    traffic.setdefault(name, []).append(rd + wr)
open(sys.argv[2], "w").write("\n".join(out) + "\n")
if len(sys.argv) > 3:
    m = {"stream_ew_kernel<0, 2>": "Stream_COPY", "stream_ew_kernel<1, 2>": "Stream_MUL", "stream_ew_kernel<2, 2>": "Stream_ADD",
         "stream_ew_kernel<3, 2>": "Stream_TRIAD", "reduce_kernel<2, 2>": "Stream_DOT", "reduce_kernel<1, 8>": "Algorithm_REDUCE_SUM",
         "scan_kernel<4>": "Algorithm_SCAN", "scan_tma_kernel<1, 0>": "Algorithm_SCAN"}
    m.update({k: "Apps_MASS3DPA" for k in traffic if k.startswith("mass3dpa_kernel")})
    m.update({k: "Apps_DIFFUSION3DPA" for k in traffic if k.startswith("diffusion3dpa_kernel")})
    m.update({k: "Apps_CONVECTION3DPA" for k in traffic if k.startswith("convection3dpa_kernel")})
    m.update({k: "Apps_LTIMES" for k in traffic if k.startswith("ltimes_dmma")})
    js = {m[k]: {"dram_bytes_per_launch": v[0], "kernel": k} for k, v in traffic.items() if k in m}
    try:                      # merge into an existing file: later captures refresh single kernels
        old = json.load(open(sys.argv[3]))
    except Exception:
        old = {}
    old.update(js)
    json.dump(old, open(sys.argv[3], "w"), indent=1)

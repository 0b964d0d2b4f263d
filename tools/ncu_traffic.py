#!/usr/bin/env python
"""profiles/r02_ncu_traffic.json from one ncu --set full capture of tools/prof_kernels.py (run here, no GPU):

    python tools/ncu_traffic.py <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <manifest.json from prof_kernels.py --manifest>
                                [profiles/r02_ncu_traffic.json]

Per suite kernel: the demangled name of the kernel instantiation that ran (the dominant launch of that suite kernel), its grid
and block, DRAM bytes read + written per launch (dram__bytes_read.sum + dram__bytes_write.sum), the ncu duration, the size and
launch shape the capture was made with (bench.py refuses the entry for any other), the commit and the library hash."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, manifest_path = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
man = json.load(open(manifest_path))

# suite kernel -> regex of the kernel function(s) it launches; the entry is the launch with the most DRAM traffic
PAT = {
    "Stream_COPY": r"stream_ew_kernel<\(?int\)?0,|stream_ew_kernel<0,", "Stream_MUL": r"stream_ew_kernel<\(?int\)?1,|stream_ew_kernel<1,",
    "Stream_ADD": r"stream_ew_kernel<\(?int\)?2,|stream_ew_kernel<2,", "Stream_TRIAD": r"stream_ew_kernel<\(?int\)?3,|stream_ew_kernel<3,",
    "Stream_DOT": r"reduce_kernel<\(?int\)?2,|reduce_kernel<2,", "Algorithm_REDUCE_SUM": r"reduce_kernel<\(?int\)?1,|reduce_kernel<1,",
    "Algorithm_SCAN": r"scan_tma_kernel|scan_kernel", "Algorithm_SORT": r"sort_onesweep_kernel<\(?bool\)?(0|false)",
    "Apps_MASS3DPA": r"mass3dpa_kernel", "Apps_DIFFUSION3DPA": r"diffusion3dpa_kernel", "Apps_CONVECTION3DPA": r"convection3dpa_kernel",
    "Apps_LTIMES": r"ltimes_", "Comm_HALO_PACKING_FUSED": r"halo_items_kernel<\(?bool\)?(0|false)",
    "Comm_HALO_EXCHANGE_FUSED": r"halo_items_kernel<\(?bool\)?(1|true)|halo_kernel", "Polybench_GEMM": r"gemm_dmma_kernel",
    "Basic_INDEXLIST": r"indexlist_tma_kernel|indexlist_kernel",
}


def scale(v, unit):
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(unit, 1.0)


try:
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except Exception:
    commit = "?"
out = {}
for kernel, meta in man["kernels"].items():
    if kernel not in PAT:
        continue
    best = None
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if not re.search(PAT[kernel], name):
            continue
        rd = scale(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = scale(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = scale(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        if best is None or rd + wr > best["dram_bytes_per_launch"]:
            best = {"kernel_name": re.sub(r"\(.*", "", name.replace("void ", "").replace("<unnamed>::", "")),
                    "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]],
                    "registers_per_thread": r[col["launch__registers_per_thread"]],
                    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "ncu_duration_us": us}
    if best:
        best.update(n=meta["n"], tuning=meta["tuning"], commit=commit, library_sha256=man.get("library_sha256"),
                    source=os.path.basename(raw))
        out[kernel] = best
try:
    old = json.load(open(out_path))
except Exception:
    old = {}
old.update(out)
json.dump(old, open(out_path, "w"), indent=1)
print(f"{len(out)} kernels -> {out_path}")
for k, v in out.items():
    print(f"  {k:28s} {v['kernel_name'][:60]:60s} {v['dram_bytes_per_launch'] / 1e6:10.1f} MB  {v['ncu_duration_us']:9.1f} us")

#!/usr/bin/env python
"""SURVEY 8f row 4, the incumbent column: the UNMODIFIED reference suite's GPU variants (Base_CUDA, RAJA_CUDA and their
cub tunings, built for sm_100 from /root/reference by oracle/build_ref_cuda.sh -> oracle/_ref/raja-perf-cuda.exe)
timed on the same B200, in the same call, with the same flags and sizes as this repo's harness
(rajaperf_b200/suite/raja-perf-b200.exe -v Base_B200).  Both binaries time a kernel the reference's way -- host timer
around the rep loop with a device synchronisation on both sides (KernelBase.hpp:271-294) -- and print the total seconds
of the rep loop in RAJAPerf-timing-Average.csv; this script divides by Reps and takes, per kernel, the FASTEST
variant-tuning of the reference as the incumbent.  Checksums of both runs are compared as well.

    python tools/incumbent_suite.py [--quick] [--out gpurun_out/r01_incumbent_suite]

Sizes are >= 4x the 126 MB L2 per array but below the BASELINE sizes, because the reference initialises every array
on one host thread (and checksums it in long double on one host thread), which costs box time, not kernel time;
this repo's numbers at the BASELINE sizes are in bench.py's `kernels` object.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "raja-perf-cuda.exe")
REF_MPI = os.path.join(ROOT, "oracle", "_ref", "raja-perf-cuda-mpi1.exe")   # CUDA + MPI code paths over oracle/mpi_stub, one rank
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "raja-perf.exe")      # the CPU-only build of the same sources
OURS = os.path.join(ROOT, "rajaperf_b200", "suite", "raja-perf-b200.exe")
CPU_REPS = 3            # reps of the Base_OpenMP / RAJA_OpenMP leg (a bounded sample: the CPU is ~50x slower)

# (tag, kernels, --size, --checkrun reps, extra flags for both, reference-only flags)
GROUPS = [
    ("stream", ["Stream_COPY", "Stream_MUL", "Stream_ADD", "Stream_TRIAD", "Stream_DOT"], 1 << 26, 20, [], []),
    ("algo", ["Algorithm_REDUCE_SUM", "Algorithm_SCAN", "Basic_INDEXLIST", "Basic_INDEXLIST_3LOOP"], 1 << 26, 20, [], []),
    ("sort", ["Algorithm_SORT", "Algorithm_SORTPAIRS"], 1 << 25, 3, [], []),
    ("mass", ["Apps_MASS3DPA"], 62500000, 10, [], []),
    ("pa", ["Apps_DIFFUSION3DPA", "Apps_CONVECTION3DPA"], 32000000, 10, [], []),
    ("ltimes", ["Apps_LTIMES"], 128000000, 10, [], []),
    ("comm", ["Comm_HALO_PACKING", "Comm_HALO_PACKING_FUSED"], 1 << 24, 50, [], []),
    ("gemm", ["Polybench_GEMM"], 1000000, 10, [], []),
    # the MPI-only kernels: the reference binary built against the MPI stand-in (oracle/build_ref_cuda_mpi.sh), ONE rank --
    # every message is a self-send through its pinned host buffers; Base_B200 runs the same 1 x 1 x 1 rank grid
    ("exchange", ["Comm_HALO_EXCHANGE", "Comm_HALO_EXCHANGE_FUSED", "Comm_HALO_SENDRECV"], 1 << 24, 50, [], []),
]
# --fast: sizes still > the 126 MB L2 per dominant array, reference = Base_CUDA only (RAJA_CUDA where no Base_CUDA exists),
# no OpenMP leg, Base_B200 through --graph only (plus the plain rep loop for the launch-bound Comm kernels)
FAST = {"stream": 1 << 25, "algo": 1 << 25, "sort": 1 << 24, "mass": 31250000, "pa": 16000000, "ltimes": 64000000,
        "gemm": 1000000, "comm": 1 << 24, "exchange": 1 << 24}
QUICK = {"stream": 1 << 24, "algo": 1 << 24, "sort": 1 << 22, "mass": 12500000, "pa": 6400000, "ltimes": 25600000,
         "gemm": 1000000, "comm": 1 << 21, "exchange": 1 << 21}


def run(cmd, log, timeout, env=None):
    t0 = time.time()
    with open(log, "w") as f:
        try:
            rc = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=timeout, env=env).returncode
        except subprocess.TimeoutExpired:
            rc = -9
    return rc, time.time() - t0


def read_timing(path):
    """-> {kernel: {"Variant-tuning": seconds for all reps}}; handles the reference's three header rows
    (title / variants / tunings) and this repo's two (title / "Variant-tuning")."""
    if not os.path.exists(path):
        return {}
    rows = [[c.strip() for c in l.rstrip("\n").split(",")] for l in open(path) if l.strip()]
    out = {}
    if len(rows) >= 3 and rows[1][0] == "Kernel" and rows[2][0] == "Kernel":
        names = [f"{v}-{t}" for v, t in zip(rows[1][1:], rows[2][1:])]
        body = rows[3:]
    else:
        names = rows[1][1:]
        body = rows[2:]
    for r in body:
        d = {}
        for name, cell in zip(names, r[1:]):
            try:
                d[name] = float(cell)
            except ValueError:
                pass                                   # "Not run"
        out[r[0]] = d
    return out


def read_reps(path):
    """RAJAPerf-kernels.csv -> {kernel: (problem size, reps, bytes/rep, flops/rep)}"""
    out = {}
    if not os.path.exists(path):
        return out
    lines = [l for l in open(path) if l.strip()]
    hdr = None
    for l in lines:
        c = [x.strip() for x in l.split(",")]
        if c[0] == "Kernels":
            hdr = c
            continue
        if hdr and len(c) >= 7:
            ix = {h: i for i, h in enumerate(hdr)}
            out[c[0]] = (int(float(c[ix["Problem size"]])), int(float(c[ix["Reps"]])), float(c[ix["Bytes/rep"]]),
                         float(c[ix["FLOPS/rep"]]))
    return out


def read_checksums(path):
    """RAJAPerf-checksum.txt -> {kernel: {"Variant-tuning": checksum string}}"""
    out, cur = {}, None
    if not os.path.exists(path):
        return out
    for l in open(path):
        t = l.split()
        if len(t) == 1 and re.match(r"^[A-Za-z]+_[A-Za-z0-9_]+$", t[0]) and not t[0].startswith(("Base_", "RAJA_")):
            cur = t[0]
            out[cur] = {}
        elif cur and len(t) >= 2 and t[0].startswith(("Base_", "RAJA_")):
            out[cur][t[0]] = t[1]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="small sizes (a functional check of the tool)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r01_incumbent_suite"))
    ap.add_argument("--groups", nargs="*", default=None)
    ap.add_argument("--timeout", type=int, default=120, help="seconds per binary per group")
    ap.add_argument("--budget", type=float, default=1e9, help="stop starting new runs after this many seconds in total")
    ap.add_argument("--no-cpu", action="store_true", help="skip the Base_OpenMP / RAJA_OpenMP leg")
    ap.add_argument("--fast", action="store_true", help="smaller sizes, Base_CUDA only, no OpenMP leg (bounded box time)")
    a = ap.parse_args()
    for exe in (REF, OURS):
        if not os.path.exists(exe):
            sys.exit(f"{exe} is missing (oracle/build_ref_cuda.sh builds the reference; __graft_entry__.build() the harness)")
    os.makedirs(a.out, exist_ok=True)
    table, wall = [], {}
    t_begin = time.time()
    threads = os.cpu_count() or 1
    omp_env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="cores")
    cpu_exe = REF_CPU if os.path.exists(REF_CPU) else REF
    order = [g for t in a.groups for g in GROUPS if g[0] == t] if a.groups else GROUPS     # --groups also orders the runs
    for tag, kernels, size, reps, both, ref_only in order:
        if time.time() - t_begin > a.budget:
            wall[tag] = {"skipped": "time budget"}
            continue
        if a.quick:
            size = QUICK[tag]
        if a.fast:
            size = FAST[tag]
            a.no_cpu = True
        ref_variants = ["Base_CUDA", "RAJA_CUDA"] if (not a.fast or tag == "sort") else ["Base_CUDA"]
        common = ["-k"] + kernels + ["--size", str(size), "--checkrun", str(reps)] + both
        rdir, odir, gdir = (os.path.join(a.out, f"{tag}_{w}") for w in ("ref", "b200", "b200_graph"))
        ref_exe = REF_MPI if tag == "exchange" else REF
        if not os.path.exists(ref_exe):
            wall[tag] = {"skipped": f"{os.path.basename(ref_exe)} is not built"}
            continue
        rc_r, t_r = run([ref_exe] + common + ["-v"] + ref_variants + ref_only + ["--outdir", rdir], rdir + ".log", a.timeout)
        if a.fast and tag not in ("comm", "exchange"):
            rc_o, t_o = None, 0.0
        else:
            rc_o, t_o = run([OURS] + common + ["-v", "Base_B200", "-t", "default", "--outdir", odir], odir + ".log", a.timeout)
        rc_g, t_g = run([OURS] + common + ["-v", "Base_B200", "-t", "default", "--graph", "--outdir", gdir], gdir + ".log", a.timeout)
        wall[tag] = {"reference_s": round(t_r, 1), "b200_s": round(t_o, 1), "b200_graph_s": round(t_g, 1), "rc": [rc_r, rc_o, rc_g]}
        cdir = os.path.join(a.out, f"{tag}_omp")
        if not a.no_cpu and tag != "exchange" and time.time() - t_begin < a.budget:
            rc_c, t_c = run([cpu_exe, "-k"] + kernels + ["--size", str(size), "--checkrun", str(CPU_REPS)] + both +
                            ["-v", "Base_OpenMP", "RAJA_OpenMP", "--outdir", cdir], cdir + ".log", a.timeout, env=omp_env)
            wall[tag].update(openmp_s=round(t_c, 1), openmp_rc=rc_c)
        cpu_t = read_timing(os.path.join(cdir, "RAJAPerf-timing-Average.csv"))
        cpu_k = read_reps(os.path.join(cdir, "RAJAPerf-kernels.csv"))
        ref_t = read_timing(os.path.join(rdir, "RAJAPerf-timing-Average.csv"))
        our_t = read_timing(os.path.join(odir, "RAJAPerf-timing-Average.csv"))
        gra_t = read_timing(os.path.join(gdir, "RAJAPerf-timing-Average.csv"))
        ref_k = read_reps(os.path.join(rdir, "RAJAPerf-kernels.csv"))
        our_k = read_reps(os.path.join(odir, "RAJAPerf-kernels.csv")) or read_reps(os.path.join(gdir, "RAJAPerf-kernels.csv"))
        ref_c = read_checksums(os.path.join(rdir, "RAJAPerf-checksum.txt"))
        our_c = read_checksums(os.path.join(odir, "RAJAPerf-checksum.txt")) or read_checksums(os.path.join(gdir, "RAJAPerf-checksum.txt"))
        for k in kernels:
            row = {"kernel": k, "size": size}
            if k in ref_t and ref_t[k] and k in ref_k:
                n, r, b, fl = ref_k[k]
                per = {name: s / r * 1e3 for name, s in ref_t[k].items() if s > 0}
                if per:
                    best = min(per, key=per.get)
                    row.update(problem_size=n, reps=r, ref_bytes_per_rep=b, incumbent=best, incumbent_ms=per[best],
                               incumbent_gbs=b / per[best] / 1e6, reference_all_ms=per)
                    row["incumbent_checksum"] = ref_c.get(k, {}).get(best)
            if k in cpu_t and cpu_t[k] and k in cpu_k:
                n, r, b, fl = cpu_k[k]
                per = {name: sec / r * 1e3 for name, sec in cpu_t[k].items() if sec > 0}
                if per:
                    best = min(per, key=per.get)
                    row.update(openmp=best, openmp_ms=per[best], openmp_gbs=b / per[best] / 1e6, openmp_threads=threads,
                               openmp_reps=r)
            for t, key in ((our_t, "b200_ms"), (gra_t, "b200_graph_ms")):
                if k in t and t[k] and k in our_k:
                    n, r, b, fl = our_k[k]
                    s = next(iter(t[k].values()))
                    row[key] = s / r * 1e3
                    row["b200_bytes_per_rep"] = b
            if "b200_ms" in row or "b200_graph_ms" in row:
                row["b200_gbs"] = row["b200_bytes_per_rep"] / row.get("b200_graph_ms", row.get("b200_ms")) / 1e6
                row["b200_checksum"] = next(iter(our_c.get(k, {}).values()), None)
            if "incumbent_ms" in row and "b200_ms" in row:
                row["speedup"] = row["incumbent_ms"] / row["b200_ms"]
            if "incumbent_ms" in row and "b200_graph_ms" in row:
                row["speedup_graph"] = row["incumbent_ms"] / row["b200_graph_ms"]
            table.append(row)
    res = {"what": "reference Base_CUDA/RAJA_CUDA (fastest tuning per kernel) vs Base_B200, same B200, same flags, host timers "
                   "around the rep loop as the reference times them; openmp = the reference's Base_OpenMP/RAJA_OpenMP (fastest) "
                   f"on the box's {threads} host threads, {CPU_REPS} reps", "wall_seconds": wall, "rows": table}
    with open(a.out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    print(f"{'kernel':26s} {'size':>11s} {'OpenMP ms':>10s} {'incumbent (fastest)':34s} {'ms/rep':>9s} {'Base_B200':>9s} {'graph':>9s} {'x':>6s} {'x graph':>7s}  checksums")
    for r in table:
        same = ""
        if r.get("incumbent_checksum") and r.get("b200_checksum"):
            x, y = float(r["incumbent_checksum"]), float(r["b200_checksum"])
            same = "equal" if r["incumbent_checksum"] == r["b200_checksum"] else f"diff {abs(x - y):.3g}"
        f = lambda key, w: f"{r[key]:{w}.4f}" if key in r else " " * (w - 1) + "-"
        g = lambda key, w: f"{r[key]:{w}.2f}" if key in r else " " * (w - 1) + "-"
        print(f"{r['kernel']:26s} {r['size']:11d} {f('openmp_ms', 10)} {r.get('incumbent', '-'):34s} {f('incumbent_ms', 9)} {f('b200_ms', 9)} "
              f"{f('b200_graph_ms', 9)} {g('speedup', 6)} {g('speedup_graph', 7)}  {same}")
    print("wall seconds per group:", json.dumps(wall))


if __name__ == "__main__":
    main()

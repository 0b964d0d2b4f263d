#!/usr/bin/env python
"""The reference adjudicates: run the reference's OWN driver with the Base_B200 variant integrated
(oracle/_ref/raja-perf-with-b200.exe, built by rajaperf_b200/integration/build_ref_b200.sh) for all 15 hot-path kernels and
keep its checksum report (common/Executor.cpp:1281-1485: every variant against the first one listed, i.e. Base_Seq) and its
timing report.  Three phases, each optional:

  default    one run, default sizes, `-v Base_Seq Base_CUDA RAJA_CUDA Base_B200 --checkrun 3`, every tuning
  checksum   BASELINE sizes, `-v Base_Seq Base_B200 --checkrun 2`, one process per kernel, a few at a time (the time goes to the
             reference's single-threaded initialisation and long-double checksum, not to kernels; timings of this phase are
             NOT used)
  timing     BASELINE sizes, sequential, `-v <incumbent variant> Base_B200 -t <incumbent tuning> default --checkrun R`:
             same-size, same-run incumbent numbers.  The incumbent (variant, tuning) per kernel is the fastest one of the
             reference's full tuning list as measured in round 1 (profiles/r01_incumbent_suite_h.md).

    python tools/ref_adjudicate.py --phases default checksum timing --out gpurun_out/r02_adjudicate

Writes <out>/<phase>[_<kernel>]/RAJAPerf-*.{txt,csv} (the reference's own report files), <out>/summary.json and summary.md.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from incumbent_suite import read_checksums, read_reps, read_timing  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "raja-perf-with-b200.exe")

# kernel -> (--size at the BASELINE configuration [SURVEY 8d], reps for the timing phase, incumbent variant, incumbent tuning)
KERNELS = {
    # (reps: the timing phase runs GPU variants only, so reps are cheap; a rep batch of a few ms sits inside the clock ramp
    #  that follows the reference's 20-30 s of host-side initialisation, hence >= 50 ms of kernel time per batch)
    "Stream_COPY": (1 << 28, 100, "RAJA_CUDA", "block_256"),
    "Stream_MUL": (1 << 28, 100, "RAJA_CUDA", "block_256"),
    "Stream_ADD": (1 << 28, 100, "RAJA_CUDA", "block_256"),
    "Stream_TRIAD": (1 << 28, 100, "RAJA_CUDA", "block_256"),
    "Stream_DOT": (1 << 28, 100, "RAJA_CUDA", "blkdev_occgs_new_256"),
    "Algorithm_REDUCE_SUM": (1 << 27, 400, "RAJA_CUDA", "blkdev_occgs_256"),
    "Algorithm_SCAN": (1 << 27, 100, "RAJA_CUDA", "cub"),
    "Algorithm_SORT": (1 << 27, 3, "RAJA_CUDA", "default"),
    "Algorithm_SORTPAIRS": (1 << 27, 3, "RAJA_CUDA", "default"),
    "Apps_MASS3DPA": (500000000, 20, "Base_CUDA", "block_25"),
    "Apps_DIFFUSION3DPA": (256000000, 20, "Base_CUDA", "block_64"),
    "Apps_CONVECTION3DPA": (256000000, 20, "Base_CUDA", "block_64"),
    "Apps_LTIMES": (1024000000, 20, "Base_CUDA", "block_256"),
    "Comm_HALO_PACKING_FUSED": (1 << 27, 200, "Base_CUDA", "direct_1024"),
}
CHECK_REPS = {"Algorithm_SORT": 2, "Algorithm_SORTPAIRS": 2, "Apps_LTIMES": 1}   # Base_Seq at these sizes: 12-25 s per rep
# The reference's own LTIMES GPU variants launch a 3-D grid with gridDim.z = num_z (LTIMES-Cuda.cpp:44-102): beyond
# num_z = 65535 the launch fails ("invalid configuration argument", r02_a / r02_b logs), so the incumbent cannot run the
# BASELINE size (num_z = 500000).  Its timing phase uses the largest size the reference can launch: num_z = 65535.
TIMING_SIZE = {"Apps_LTIMES": 65535 * 2048}


def run(cmd, log, timeout):
    t0 = time.time()
    with open(log, "w") as f:
        try:
            rc = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=timeout).returncode
        except subprocess.TimeoutExpired:
            rc = -9
    return rc, round(time.time() - t0, 1)


def mem_available_gb():
    for l in open("/proc/meminfo"):
        if l.startswith("MemAvailable"):
            return int(l.split()[1]) / 1e6
    return 0.0


def diffs(cks):
    """{variant-tuning: checksum string} -> {variant-tuning: abs diff against Base_Seq-default} (decimal arithmetic)"""
    from decimal import Decimal
    ref = next((v for k, v in cks.items() if k.startswith("Base_Seq-")), None)     # "-default", or "-direct" for the Comm kernels
    if ref is None:
        return {}
    return {k: float(abs(Decimal(v) - Decimal(ref))) for k, v in cks.items()}


def collect(odir, kernels):
    t = read_timing(os.path.join(odir, "RAJAPerf-timing-Average.csv"))
    k = read_reps(os.path.join(odir, "RAJAPerf-kernels.csv"))
    c = read_checksums(os.path.join(odir, "RAJAPerf-checksum.txt"))
    rows = {}
    for name in kernels:
        row = {}
        if name in c:
            row["checksums"] = c[name]
            row["abs_diff_vs_Base_Seq"] = diffs(c[name])
        if name in t and name in k:
            n, reps, nbytes, flops = k[name]
            row.update(problem_size=n, reps=reps, bytes_per_rep=nbytes, flops_per_rep=flops,
                       ms_per_rep={v: s / reps * 1e3 for v, s in t[name].items() if s > 0})
        rows[name] = row
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phases", nargs="+", default=["default", "checksum", "timing"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_adjudicate"))
    ap.add_argument("--kernels", nargs="*", default=None)
    ap.add_argument("--jobs", type=int, default=5, help="processes at a time in the checksum phase")
    ap.add_argument("--timeout", type=int, default=420, help="seconds per process")
    ap.add_argument("--quick", action="store_true", help="1/64 of the BASELINE sizes (a functional check of the tool)")
    a = ap.parse_args()
    if not os.path.exists(EXE):
        sys.exit(f"{EXE} is missing (rajaperf_b200/integration/build_ref_b200.sh builds it)")
    os.makedirs(a.out, exist_ok=True)
    names = [k for k in KERNELS if not a.kernels or k in a.kernels]
    summary = {"host_threads": os.cpu_count(), "mem_available_gb": round(mem_available_gb(), 1), "phases": {}}
    size = lambda k: KERNELS[k][0] // (64 if a.quick else 1)

    if "default" in a.phases:
        odir = os.path.join(a.out, "default")
        rc, sec = run([EXE, "-k"] + names + ["-v", "Base_Seq", "Base_CUDA", "RAJA_CUDA", "Base_B200", "--checkrun", "3",
                                             "--outdir", odir], odir + ".log", a.timeout)
        summary["phases"]["default"] = {"rc": rc, "wall_s": sec, "kernels": collect(odir, names)}

    if "checksum" in a.phases:
        jobs = a.jobs if mem_available_gb() > 96 else 2

        def one(k):
            odir = os.path.join(a.out, "checksum_" + k)
            rc, sec = run([EXE, "-k", k, "-v", "Base_Seq", "Base_B200", "--size", str(size(k)), "--checkrun",
                           str(CHECK_REPS.get(k, 2)), "--outdir", odir], odir + ".log", a.timeout)
            return k, rc, sec, collect(odir, [k])[k]
        t0 = time.time()
        res = {}
        # the slowest first
        order = sorted(names, key=lambda k: -(("SORT" in k) * 4 + ("LTIMES" in k) * 3 + ("3DPA" in k) * 2 + ("Stream" in k)))
        with cf.ThreadPoolExecutor(jobs) as ex:
            for k, rc, sec, row in ex.map(one, order):
                row.update(rc=rc, wall_s=sec)
                res[k] = row
        summary["phases"]["checksum"] = {"jobs": jobs, "wall_s": round(time.time() - t0, 1), "kernels": {k: res[k] for k in names}}

    if "timing" in a.phases:
        res = {}
        t0 = time.time()
        for k in names:
            _, reps, var, tune = KERNELS[k]
            odir = os.path.join(a.out, "timing_" + k)
            tsize = TIMING_SIZE.get(k, KERNELS[k][0]) // (64 if a.quick else 1)
            rc, sec = run([EXE, "-k", k, "-v", var, "Base_B200", "-t", tune, "default", "--size", str(tsize), "--checkrun",
                           str(reps), "--outdir", odir], odir + ".log", a.timeout)
            row = collect(odir, [k])[k]
            row.update(rc=rc, wall_s=sec, incumbent=f"{var}-{tune}")
            ms = row.get("ms_per_rep", {})
            inc = {v: t for v, t in ms.items() if not v.startswith("Base_B200")}
            ours = ms.get("Base_B200-default")
            if inc and ours:
                best = min(inc, key=inc.get)
                row.update(incumbent=best, incumbent_ms=inc[best], b200_ms=ours, speedup=inc[best] / ours,
                           incumbent_gbs=row["bytes_per_rep"] / inc[best] / 1e6, b200_gbs=row["bytes_per_rep"] / ours / 1e6)
            res[k] = row
        summary["phases"]["timing"] = {"wall_s": round(time.time() - t0, 1), "kernels": res}

    json.dump(summary, open(os.path.join(a.out, "summary.json"), "w"), indent=1)
    # ---- markdown
    L = ["# The reference driver adjudicates `Base_B200` (tools/ref_adjudicate.py)", "",
         f"host threads {summary['host_threads']}, MemAvailable {summary['mem_available_gb']} GB", ""]
    for ph in ("default", "checksum"):
        if ph not in summary["phases"]:
            continue
        P = summary["phases"][ph]
        L += [f"## phase `{ph}` ({P['wall_s']} s)", "", "| kernel | problem size | reps | Base_Seq checksum | variant-tuning: abs diff vs Base_Seq |",
              "|---|---|---|---|---|"]
        for k, row in P["kernels"].items():
            d = row.get("abs_diff_vs_Base_Seq", {})
            cells = "; ".join(f"`{v}` {x:.3g}" for v, x in d.items() if not v.startswith("Base_Seq-") and
                              (ph == "checksum" or v.startswith("Base_B200") or x > 0))
            L.append(f"| {k} | {row.get('problem_size', '')} | {row.get('reps', '')} | "
                     f"{next((v for n_, v in row.get('checksums', {}).items() if n_.startswith('Base_Seq-')), 'MISSING')} | {cells or 'all variants 0'} |")
        L.append("")
    if "timing" in summary["phases"]:
        P = summary["phases"]["timing"]
        L += [f"## phase `timing` ({P['wall_s']} s): same size, same run, the reference's own timers", "",
              "| kernel | problem size | reps | incumbent | incumbent ms/rep | GB/s | Base_B200 ms/rep | GB/s | speed-up | checksums equal |",
              "|---|---|---|---|---|---|---|---|---|---|"]
        for k, row in P["kernels"].items():
            if "speedup" not in row:
                L.append(f"| {k} | rc {row.get('rc')} | | | | | | | | |")
                continue
            c = row.get("checksums", {})
            eq = len(set(c.values())) == 1
            L.append(f"| {k} | {row['problem_size']} | {row['reps']} | `{row['incumbent']}` | {row['incumbent_ms']:.4f} | "
                     f"{row['incumbent_gbs']:.0f} | {row['b200_ms']:.4f} | {row['b200_gbs']:.0f} | {row['speedup']:.2f}x | "
                     f"{'yes' if eq else '; '.join(f'{v} {x}' for v, x in c.items())} |")
    open(os.path.join(a.out, "summary.md"), "w").write("\n".join(L) + "\n")
    print("\n".join(L))


if __name__ == "__main__":
    main()

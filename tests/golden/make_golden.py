#!/usr/bin/env python
"""Mint known-answer checksums from the UNMODIFIED reference binary (Base_Seq variant).

Usage (in the authoring container, where /root/reference exists):
    python tests/golden/make_golden.py [--exe oracle/_ref/raja-perf.exe]

The binary is the reference's own raja-perf.exe built CPU-only from /root/reference
(oracle/build_ref.sh: the reference's own CMake build, CPU-only, out of tree).  Every case runs
    raja-perf.exe --checkrun R --disable-warmup -k K -v Base_Seq [--size S] [kernel flags]
and the Base_Seq checksum printed in RAJAPerf-checksum.txt is recorded verbatim (20 digits).
Output: tests/golden/ref_checksums.json.  The GPU box never runs this script.

    python tests/golden/make_golden.py --mpi      (-> tests/golden/ref_checksums_mpi1.json)
mints the goldens of the three kernels the reference only compiles WITH MPI -- Comm_HALO_EXCHANGE,
Comm_HALO_EXCHANGE_FUSED, Comm_HALO_SENDRECV -- from oracle/_ref/raja-perf-mpi1.exe: the same unmodified sources built
with ENABLE_MPI=On against the one-rank in-process MPI stand-in of oracle/mpi_stub/ (oracle/build_ref_mpi.sh), i.e. the
periodic self-exchange of a 1 x 1 x 1 rank grid.  The HALO_PACKING / HALO_PACKING_FUSED cases are re-run there too: they
must (and do) equal the non-MPI binary's.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (kernel, size or 0 for default, reps, extra flags)
CASES = []
for k in ["Stream_COPY", "Stream_MUL", "Stream_ADD", "Stream_TRIAD", "Stream_DOT",
          "Algorithm_REDUCE_SUM", "Algorithm_SCAN", "Algorithm_SORT", "Algorithm_SORTPAIRS"]:
    CASES += [(k, 0, 1, []), (k, 0, 3, []), (k, 1, 1, []), (k, 1000, 2, []), (k, 123457, 2, [])]
for k in ["Apps_MASS3DPA", "Apps_DIFFUSION3DPA", "Apps_CONVECTION3DPA"]:
    CASES += [(k, 0, 1, []), (k, 0, 3, []), (k, 1, 1, []), (k, 5000, 2, [])]
CASES += [("Apps_LTIMES", 0, 1, []), ("Apps_LTIMES", 0, 3, []), ("Apps_LTIMES", 5000, 2, []),
          ("Apps_LTIMES", 100000, 1, ["--ltimes_num_d", "32", "--ltimes_num_g", "8", "--ltimes_num_m", "17"])]
CASES += [("Comm_HALO_PACKING_FUSED", 0, 1, []), ("Comm_HALO_PACKING_FUSED", 0, 3, []),
          ("Comm_HALO_PACKING_FUSED", 27000, 2, ["--halo_width", "2", "--halo_num_vars", "5"]),
          ("Comm_HALO_PACKING_FUSED", 1000, 1, ["--halo_width", "3", "--halo_num_vars", "1"])]
# widened rows
for k in ["Basic_INDEXLIST", "Basic_INDEXLIST_3LOOP"]:
    CASES += [(k, 0, 1, []), (k, 0, 3, []), (k, 1, 1, []), (k, 1000, 2, []), (k, 123457, 2, [])]
CASES += [("Polybench_GEMM", 0, 1, []), ("Polybench_GEMM", 0, 2, []), ("Polybench_GEMM", 1, 1, []),
          ("Polybench_GEMM", 10000, 2, []), ("Polybench_GEMM", 54321, 1, [])]
for k in ["Algorithm_MEMCPY", "Algorithm_MEMSET"]:      # calibration streams: the checksum is 0 iff every element was written
    CASES += [(k, 0, 1, []), (k, 1, 1, []), (k, 123457, 2, [])]
CASES += [("Comm_HALO_PACKING", 0, 1, []), ("Comm_HALO_PACKING", 27000, 2, ["--halo_width", "2", "--halo_num_vars", "5"])]


# the kernels that exist only in an MPI build: (kernel, size, reps, flags, ranks).  One rank = the in-process transport of
# the stub (+ the two pack kernels as a cross-check of the binary); P ranks = P processes of the same binary over the stub's
# shared-memory transport, rank grid given explicitly or left to the suite's default factorisation (RunParams.cpp:1211-1251)
MPI_CASES = []
for k in ["Comm_HALO_EXCHANGE_FUSED", "Comm_HALO_EXCHANGE", "Comm_HALO_SENDRECV"]:
    MPI_CASES += [(k, 0, 1, [], 1), (k, 0, 3, [], 1), (k, 27000, 2, ["--halo_width", "2", "--halo_num_vars", "5"], 1),
                  (k, 1000, 1, ["--halo_width", "3", "--halo_num_vars", "1"], 1), (k, 8000, 2, [], 1),
                  (k, 27000, 2, ["--halo_width", "2", "--halo_num_vars", "2"], 1)]
    MPI_CASES += [(k, 27000, 2, [], 2), (k, 8000, 3, ["--halo_width", "2", "--halo_num_vars", "2"], 4),
                  (k, 27000, 2, ["--halo_width", "2", "--halo_num_vars", "2", "--mpi_3d_division", "2", "2", "2"], 8),
                  (k, 8000, 2, ["--mpi_3d_division", "3", "1", "2"], 6),
                  (k, 1000, 1, ["--halo_width", "3", "--halo_num_vars", "1", "--mpi_3d_division", "1", "2", "1"], 2)]
MPI_CASES += [("Comm_HALO_PACKING_FUSED", 0, 1, [], 1), ("Comm_HALO_PACKING_FUSED", 27000, 2, ["--halo_width", "2", "--halo_num_vars", "5"], 1),
              ("Comm_HALO_PACKING", 0, 1, [], 1)]


def fuzz_cases(n_per_kernel, seed):
    """Seeded random (size, reps, flags) draws per kernel: ragged sizes around the kernels' tile boundaries (8192-element
    tiles, 2048-element halo chunks, 8-element PA batches), odd LTIMES shapes, halo widths 1-3 with 1-6 variables, and for the
    MPI-only kernels random rank grids of 1-8 ranks.  (kernel, size, reps, flags, ranks); ranks = 0 marks a case of the
    non-MPI binary."""
    import random
    rng = random.Random(seed)

    def near(*anchors):
        a = rng.choice(anchors)
        return max(1, a + rng.choice([-3, -1, 0, 1, 2, 7]) + (rng.randrange(0, a) if rng.random() < 0.3 else 0))

    out = []
    for k in ["Stream_COPY", "Stream_MUL", "Stream_ADD", "Stream_TRIAD", "Stream_DOT", "Algorithm_REDUCE_SUM", "Algorithm_SCAN",
              "Algorithm_SORT", "Algorithm_SORTPAIRS", "Basic_INDEXLIST", "Basic_INDEXLIST_3LOOP", "Algorithm_MEMCPY", "Algorithm_MEMSET"]:
        for _ in range(n_per_kernel):
            out.append((k, near(5, 33, 257, 2048, 8192, 16384, 3 * 8192, 65536, 200000), rng.randint(1, 4), [], 0))
    for k in ["Apps_MASS3DPA", "Apps_DIFFUSION3DPA", "Apps_CONVECTION3DPA"]:
        for _ in range(n_per_kernel):
            out.append((k, near(70, 125 * 7, 125 * 8 + 60, 64 * 9, 64 * 33, 20000), rng.randint(1, 3), [], 0))
    for _ in range(n_per_kernel):
        d, g, m = rng.choice([1, 3, 8, 17, 32, 64, 70]), rng.choice([1, 2, 8, 13, 32]), rng.choice([1, 5, 17, 25, 31])
        flags = ["--ltimes_num_d", str(d), "--ltimes_num_g", str(g), "--ltimes_num_m", str(m)] if rng.random() < 0.7 else []
        out.append(("Apps_LTIMES", near(2048, 2048 * 9, 50000, 150000), rng.randint(1, 3), flags, 0))
    for _ in range(n_per_kernel):
        out.append(("Polybench_GEMM", near(1, 50, 900, 4096, 30000), rng.randint(1, 2), [], 0))

    def halo_flags():
        w, nv = rng.randint(1, 3), rng.randint(1, 6)
        side = rng.randint(2 * w, 34)                       # a box at least as wide as two halos in every direction
        return side ** 3 + rng.choice([0, 0, 1, 5]), ["--halo_width", str(w), "--halo_num_vars", str(nv)]
    for k in ["Comm_HALO_PACKING_FUSED", "Comm_HALO_PACKING"]:
        for _ in range(n_per_kernel):
            size, fl = halo_flags()
            out.append((k, size, rng.randint(1, 3), fl, 0))
    grids = [(1, 1, 1), (2, 1, 1), (1, 2, 1), (1, 1, 2), (3, 1, 1), (2, 2, 1), (2, 1, 2), (5, 1, 1), (3, 2, 1), (1, 2, 3), (7, 1, 1), (2, 2, 2), (4, 2, 1)]
    for k in ["Comm_HALO_EXCHANGE_FUSED", "Comm_HALO_EXCHANGE", "Comm_HALO_SENDRECV"]:
        for _ in range(n_per_kernel):
            size, fl = halo_flags()
            gx, gy, gz = rng.choice(grids)
            out.append((k, size, rng.randint(1, 3), fl + ["--mpi_3d_division", str(gx), str(gy), str(gz)], gx * gy * gz))
    return out


def mpirun(nranks, cmd, timeout=600):
    """P processes of `cmd` over the stub's shared-memory transport (oracle/mpi_stub/mpi_stub.c): one zero-filled arena
    under /dev/shm, RPB_MPI_SIZE / RPB_MPI_RANK / RPB_MPI_SHM in the environment; rank 0 writes the reports."""
    path = f"/dev/shm/rpb_mpi_{os.getpid()}"
    with open(path, "wb") as f:
        f.truncate(256 << 20)
    try:
        procs = [subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                                  env=dict(os.environ, OMP_NUM_THREADS="1", RPB_MPI_SIZE=str(nranks), RPB_MPI_RANK=str(r),
                                           RPB_MPI_SHM=path)) for r in range(nranks)]
        rcs = [p.wait(timeout=timeout) for p in procs]
    finally:
        os.unlink(path)
    if any(rcs):
        raise RuntimeError(f"{nranks}-rank run failed: {rcs}: {' '.join(cmd)}")


# --dryrun tables: what the reference prints for (problem size, reps, iterations/rep, kernels/rep, bytes/rep, flops/rep) of
# every hot-path kernel under these flag sets -- the kernel-class metadata the harness must reproduce (KernelBase.hpp:100-116)
DRYRUN_KERNELS = ["Stream", "Algorithm_REDUCE_SUM", "Algorithm_SCAN", "Algorithm_SORT", "Algorithm_SORTPAIRS", "Algorithm_MEMCPY",
                  "Algorithm_MEMSET", "Apps_MASS3DPA", "Apps_DIFFUSION3DPA", "Apps_CONVECTION3DPA", "Apps_LTIMES", "Comm",
                  "Basic_INDEXLIST", "Basic_INDEXLIST_3LOOP", "Polybench_GEMM"]
DRYRUN_FLAGS = [[], ["--size", "1234567"], ["--sizefact", "2.5", "--repfact", "0.3"],
                ["--size", "27000", "--halo_width", "2", "--halo_num_vars", "5"], ["--repfact", "2", "--size", "50000"],
                ["--size", "268435456"], ["--size", "100000", "--ltimes_num_d", "32", "--ltimes_num_g", "8", "--ltimes_num_m", "17"]]


def dryrun_tables(exe):
    out = []
    for flags in DRYRUN_FLAGS:
        txt = subprocess.run([exe, "--dryrun", "-k"] + DRYRUN_KERNELS + ["-v", "Base_Seq", "RAJA_Seq"] + flags,
                             capture_output=True, text=True, check=True).stdout
        rows = {}
        for line in txt.splitlines():
            if re.match(r"^(Stream|Algorithm|Apps|Comm|Basic|Polybench)_", line):
                f = [x.strip() for x in line.split(",")]
                rows[f[0]] = f[1:7]
        out.append({"flags": flags, "rows": rows})
    return out


def run_case(exe, kernel, size, reps, extra, workdir, ranks=1):
    out = os.path.join(workdir, "out")
    shutil.rmtree(out, ignore_errors=True)
    cmd = [exe, "--checkrun", str(reps), "--disable-warmup", "-k", kernel, "-v", "Base_Seq",
           "--outdir", out] + (["--size", str(size)] if size else []) + extra
    if ranks > 1:
        mpirun(ranks, cmd)
    else:
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    txt = open(os.path.join(out, "RAJAPerf-checksum.txt")).read()
    m = re.search(r"^Base_Seq-\S+\s+(\S+)", txt, re.M)
    if not m:
        raise RuntimeError(f"no Base_Seq checksum for {kernel}:\n{txt}")
    return m.group(1)


def fuzz_main(a):
    plain = os.path.join(ROOT, "oracle", "_ref", "raja-perf.exe")
    mpi = os.path.join(ROOT, "oracle", "_ref", "raja-perf-mpi1.exe")
    work = os.path.join(ROOT, "build", "golden_work")
    os.makedirs(work, exist_ok=True)
    rows = []
    for kernel, size, reps, extra, ranks in fuzz_cases(a.fuzz, a.seed):
        try:
            ck = run_case(mpi if ranks else plain, kernel, size, reps, extra, work, max(ranks, 1))
        except Exception as e:                    # the reference itself refuses or crashes on this input: not a golden
            print("SKIPPED", kernel, size, reps, extra, ranks, str(e)[:120], file=sys.stderr)
            continue
        rows.append({"kernel": kernel, "size": size, "reps": reps, "flags": extra, "variant": "Base_Seq", "checksum": ck, "ranks": ranks})
        print(kernel, size, reps, extra, ranks, ck, file=sys.stderr)
    out = a.out or os.path.join(ROOT, "tests", "golden", "ref_checksums_fuzz.json")
    json.dump({"source": "reference raja-perf.exe / raja-perf-mpi1.exe (suite v2024.07.0, commit 9af20b3; the CPU-only builds of "
                         "oracle/build_ref.sh and oracle/build_ref_mpi.sh), g++ 13.3 -O3, glibc rand(), x87 long double",
               "command": f"python tests/golden/make_golden.py --fuzz {a.fuzz} --seed {a.seed}   (ranks = 0: the non-MPI binary; "
                          "ranks >= 1: P processes of the MPI-stub binary, the checksum is the report's average over the ranks)",
               "cases": rows}, open(out, "w"), indent=1)
    print(f"wrote {len(rows)} cases to {out}", file=sys.stderr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exe", default=os.path.join(ROOT, "oracle", "_ref", "raja-perf.exe"))
    ap.add_argument("--out", default=None)
    ap.add_argument("--mpi", action="store_true", help="the MPI-only Comm kernels from the one-rank MPI-stub build")
    ap.add_argument("--dryrun", action="store_true", help="the --dryrun tables of the MPI-stub build (all 23 kernels) -> ref_dryrun.json")
    ap.add_argument("--fuzz", type=int, default=0, help="N seeded random cases per kernel from both binaries -> ref_checksums_fuzz.json")
    ap.add_argument("--seed", type=int, default=20261018)
    a = ap.parse_args()
    if a.fuzz:
        fuzz_main(a)
        return
    if a.dryrun:
        exe = os.path.join(ROOT, "oracle", "_ref", "raja-perf-mpi1.exe")
        out = a.out or os.path.join(ROOT, "tests", "golden", "ref_dryrun.json")
        json.dump({"source": "reference raja-perf.exe --dryrun (suite v2024.07.0, CPU-only build against oracle/mpi_stub, one rank)",
                   "columns": ["Problem size", "Reps", "Iterations/rep", "Kernels/rep", "Bytes/rep", "FLOPS/rep"],
                   "tables": dryrun_tables(exe)}, open(out, "w"), indent=1)
        print(f"wrote {out}", file=sys.stderr)
        return
    cases = CASES
    if a.mpi:
        cases = MPI_CASES
        if a.exe.endswith("raja-perf.exe"):
            a.exe = os.path.join(ROOT, "oracle", "_ref", "raja-perf-mpi1.exe")
    if a.out is None:
        a.out = os.path.join(ROOT, "tests", "golden", "ref_checksums_mpi1.json" if a.mpi else "ref_checksums.json")
    work = os.path.join(ROOT, "build", "golden_work")
    os.makedirs(work, exist_ok=True)
    rows = []
    for case in cases:
        kernel, size, reps, extra = case[:4]
        ranks = case[4] if len(case) > 4 else 1
        ck = run_case(a.exe, kernel, size, reps, extra, work, ranks)
        rows.append({"kernel": kernel, "size": size, "reps": reps, "flags": extra, "variant": "Base_Seq",
                     "checksum": ck})
        if a.mpi:
            rows[-1]["ranks"] = ranks        # the checksum is the report's average over the ranks
        print(kernel, size, reps, extra, ranks, ck, file=sys.stderr)
    ver = subprocess.run([a.exe, "--help"], capture_output=True, text=True).stdout.splitlines()[:1]
    json.dump({"source": ("reference raja-perf.exe (suite v2024.07.0, commit 9af20b3), CPU-only build with ENABLE_MPI=On against "
                          "the MPI stand-in oracle/mpi_stub (one rank in process, or P processes over its shared-memory transport), " if a.mpi else
                          "reference raja-perf.exe (suite v2024.07.0, commit 9af20b3), CPU-only build, ") +
                         "g++ 13.3 -O3, glibc rand(), x87 long double",
               "command": "raja-perf.exe --checkrun R --disable-warmup -k K -v Base_Seq [--size S] [flags]",
               "cases": rows}, open(a.out, "w"), indent=1)
    print(f"wrote {len(rows)} cases to {a.out}", file=sys.stderr)


if __name__ == "__main__":
    main()

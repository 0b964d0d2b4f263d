#!/usr/bin/env python
"""Launch P ranks of a binary linked against the MPI stand-in (oracle/mpi_stub): the `mpirun` of this repo's reference builds.

    python tools/mpirun_stub.py -n 8 [--gpu-per-rank] -- oracle/_ref/raja-perf-with-b200-mpi.exe \
        -k Comm_HALO_EXCHANGE_FUSED -v Base_Seq Base_CUDA Base_B200 --checkrun 5 --size 2097152

One zero-filled arena under /dev/shm carries the messages (RPB_MPI_SHM), RPB_MPI_SIZE / RPB_MPI_RANK tell a process who it
is; with --gpu-per-rank device 0 of rank r is GPU r mod <number of GPUs> (CUDA_VISIBLE_DEVICES is the GPU list rotated by r,
so the peers stay visible for CUDA IPC), the binding a real launcher would do.  Rank 0's output is shown, the others' is dropped unless --all-output.  Exit code: the first non-zero one.
"""
import argparse
import os
import subprocess
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, required=True)
    ap.add_argument("--gpu-per-rank", action="store_true")
    ap.add_argument("--all-output", action="store_true")
    ap.add_argument("--arena-mb", type=int, default=0, help="shared-memory arena (default: 512 MB per rank, at least 1024)")
    ap.add_argument("--timeout", type=int, default=900)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    if not cmd:
        ap.error("no command")
    ngpu = 0
    if a.gpu_per_rank:
        try:
            ngpu = len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
        except OSError:
            ngpu = 0
        if ngpu == 0:
            sys.exit("--gpu-per-rank: no GPU visible")
    import tempfile
    import time
    arena_mb = a.arena_mb or max(1024, 512 * a.n)
    path = f"/dev/shm/rpb_mpi_{os.getpid()}"
    with open(path, "wb") as f:
        f.truncate(arena_mb << 20)
    logs = []
    try:
        procs = []
        for r in range(a.n):
            env = dict(os.environ, RPB_MPI_SIZE=str(a.n), RPB_MPI_RANK=str(r), RPB_MPI_SHM=path)
            env.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // a.n)))
            if ngpu:      # rotated list: device 0 is this rank's GPU, the others stay visible (CUDA IPC needs the peer's device)
                env["CUDA_VISIBLE_DEVICES"] = ",".join(str((r + i) % ngpu) for i in range(ngpu))
            quiet = r != 0 and not a.all_output
            log = tempfile.TemporaryFile(mode="w+") if quiet else None       # kept: shown if the rank fails
            logs.append(log)
            procs.append(subprocess.Popen(cmd, env=env, stdout=log if quiet else None, stderr=subprocess.STDOUT if quiet else None))
        # poll: the first rank that exits non-zero takes the others down (they would wait for it until their own time-out)
        t0 = time.time()
        rcs = [None] * a.n
        while any(rc is None for rc in rcs):
            for r, p in enumerate(procs):
                if rcs[r] is None:
                    rcs[r] = p.poll()
            failed = [r for r, rc in enumerate(rcs) if rc not in (None, 0)]
            if failed or time.time() - t0 > a.timeout:
                for r, p in enumerate(procs):
                    if rcs[r] is None:
                        p.kill()
                        rcs[r] = p.wait()
                if not failed:
                    print(f"[mpirun_stub] time-out after {a.timeout} s", file=sys.stderr)
                    rcs = [rc if rc else 124 for rc in rcs]
                break
            time.sleep(0.05)
        for r, rc in enumerate(rcs):
            if rc and logs[r] is not None:
                logs[r].seek(0)
                tail = logs[r].read()[-2000:]
                print(f"[mpirun_stub] rank {r} exited with {rc}; its output ended with:\n{tail}", file=sys.stderr)
    finally:
        os.unlink(path)
    bad = [rc for rc in rcs if rc]
    sys.exit(bad[0] if bad else 0)


if __name__ == "__main__":
    main()

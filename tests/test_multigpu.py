"""Multi-rank paths.  CPU (gloo, world_size 2): the host-side plumbing in rajaperf_b200/dist.py.
GPU (>= 2 devices): tests/mgpu_check.py under torchrun -- halo exchange over NVLink peer windows and
the sharded DOT / REDUCE_SUM, both against the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_grid_follows_the_reference_factorisation():
    from rajaperf_b200.dist import rank_grid
    assert [rank_grid(p) for p in (1, 2, 3, 4, 6, 8, 12, 16)] == \
        [[1, 1, 1], [2, 1, 1], [3, 1, 1], [2, 2, 1], [2, 3, 1], [2, 2, 2], [2, 2, 3], [4, 2, 2]]
    # what the unmodified reference prints ("3D division = a x b x c") when P processes of its MPI build start over the MPI
    # stand-in:  python tools/mpirun_stub.py -n P -- oracle/_ref/raja-perf-mpi1.exe --dryrun -k Comm_HALO_EXCHANGE_FUSED
    printed = {1: [1, 1, 1], 2: [2, 1, 1], 3: [3, 1, 1], 5: [5, 1, 1], 6: [2, 3, 1], 7: [7, 1, 1], 8: [2, 2, 2], 9: [3, 3, 1],
               10: [2, 5, 1], 12: [2, 2, 3], 16: [4, 2, 2], 18: [2, 3, 3], 24: [6, 2, 2], 27: [3, 3, 3], 30: [2, 3, 5], 32: [4, 4, 2]}
    assert {p: rank_grid(p) for p in printed} == printed


def test_shard_ranges_cover_and_align():
    from rajaperf_b200.dist import shard_range
    for n in (0, 1, 5, 1000, 3000001, 1 << 28):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert b == pos and b <= e <= n and (b % 4 == 0 or b == n)
                pos = e
            assert pos == n


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from rajaperf_b200 import dist as rdist
import oracle, suite_data as sd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# 1. IPC-handle all-gather keeps rank order
h = bytes([rank]) * 64
hs = rdist.gather_handles(h)
assert [x[0] for x in hs] == list(range(world))
# 2. sharded DOT: per-shard partial (here: the oracle stands in for the kernel) + all-reduce == whole
n = 100003
d = sd.stream_dot(n)
b, e = rdist.shard_range(n, rank, world)
part = torch.tensor([oracle.lib().orc_stream_dot(d["a"][b:e].copy(), d["b"][b:e].copy(), e - b, 0.0)], dtype=torch.float64)
rdist.allreduce_scalar(part)
ref = oracle.lib().orc_stream_dot(d["a"], d["b"], n, 0.0)
assert abs(part.item() - ref) < 1e-9, (part.item(), ref)
# 3. every rank derives the same rank grid and mutually consistent neighbours
pd = rdist.rank_grid(world)
r, st, rt = sd.halo_neighbors(rank, pd)
allr = [None] * world
dist.all_gather_object(allr, (r.tolist(), st.tolist(), rt.tolist()))
for l in range(26):
    peer = allr[r[l]]
    lo = rt[l]                      # the peer's message with send tag rt[l] is addressed to me
    assert peer[0][lo] == rank and peer[1][lo] == rt[l]
dist.destroy_process_group()
print("GLOO_OK", rank)
'''


def test_gloo_world_size_2_host_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("GLOO_OK") == 2


@pytest.mark.gpu
def test_halo_exchange_and_global_reductions_across_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

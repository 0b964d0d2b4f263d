"""Host-side plumbing for the multi-GPU paths (one process per GPU, torch.distributed).

Only two hot-path pieces span GPUs: the halo exchange (ranks on a 3-D periodic grid,
windows mapped into each other through CUDA IPC) and the global DOT / REDUCE_SUM (contiguous shards,
one scalar all-reduced).  torch.distributed is used for rendezvous, the IPC-handle all-gather and the
scalar all-reduce -- plumbing; the data path is the kernels in csrc/halo.cu and csrc/stream.cu."""
from __future__ import annotations


def rank_grid(nranks: int):
    """The suite's default --mpi_3d_division (RunParams.cpp:1211-1251): prime factors in non-decreasing
    order, each multiplied into the currently smallest dimension (first one on ties)."""
    factors, number, f = [], nranks, 2
    while f * f <= number:
        if number % f == 0:
            factors.append(f)
            number //= f
        else:
            f += 1
    factors.append(number)
    dims = [1, 1, 1]
    for f in factors:
        dims[dims.index(min(dims))] *= f
    return dims


def shard_range(n: int, rank: int, world: int, align: int = 4):
    """Contiguous [begin, end) of rank's shard of an n-element array; shard starts are multiples of
    `align` elements so every shard keeps the 32-byte alignment the vector kernels want."""
    per = -(-n // world)
    per = -(-per // align) * align
    b = min(n, rank * per)
    e = min(n, b + per)
    return b, e


def gather_handles(handle: bytes, group=None):
    """All-gather the 64-byte CUDA IPC handle of this rank's halo window; returns the list by rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, handle, group=group)
    assert all(isinstance(h, (bytes, bytearray)) and len(h) == 64 for h in out)
    return [bytes(h) for h in out]


def connect_halo_plan(plan, vars_, group=None):
    """window() -> all-gather of the IPC handles -> connect(); a single rank connects to itself."""
    import torch.distributed as dist
    _, _, handle = plan.window(vars_)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        plan.connect(gather_handles(handle, group))
        dist.barrier(group)
    else:
        plan.connect_ptrs([0])


def allreduce_scalar(t, group=None):
    """Sum a 1-element tensor over ranks in rank order semantics (NCCL/gloo all_reduce SUM)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t

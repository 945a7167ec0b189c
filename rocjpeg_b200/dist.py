"""Multi-GPU plumbing for bench.py and tests: one process per GPU, no data-path collective.

rocJPEG's API binds a decoder handle to one device (src/rocjpeg_decoder.cpp:46-61) and images are
independent, so scaling = independent shards. The only cross-rank traffic is bookkeeping: a barrier
around the timed region and a MAX over per-rank times (torch.distributed; NCCL on GPUs, gloo on CPU).
"""
from __future__ import annotations

import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str, device=None):
    import torch.distributed as dist

    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return rank, world


_cpu_group = None


def cpu_barrier():
    """A barrier that keeps the GPUs idle while ranks wait (gloo over the loopback): an NCCL barrier parks a spinning
    kernel on every waiting rank's GPU, which would disturb a measurement another rank makes on those GPUs."""
    global _cpu_group
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return
    if _cpu_group is None:
        _cpu_group = dist.new_group(backend="gloo")
    dist.barrier(group=_cpu_group)


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(values, device="cpu"):
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def shard_by_cost(costs, world: int):
    """Longest-processing-time-first assignment of items (cost = entropy-coded bytes) to `world`
    ranks. Returns a list of index lists, each in ascending order; deterministic on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def finalize():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()

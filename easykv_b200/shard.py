"""Multi-GPU layout of the path: sequences are sharded across ranks, one process per GPU, and NOTHING on
the data path crosses GPUs — eviction is independent per (sequence, layer, kv head) (the reference does not
even batch: easykv/easykv.py:66-67,290,430 index `[0]`).  The only inter-rank traffic is the reduction of
timings / counts for reporting, which is why this module needs a process group at all.
"""
from __future__ import annotations

import torch


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) of `n_items` sequences for `rank` (the first n_items % world ranks get
    one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_job(elapsed_ms: float, units: int, device=None, group=None):
    """Whole-job numbers for a sharded run: (max over ranks of the device time, sum over ranks of the units
    processed).  With no initialised process group (single GPU) it is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(elapsed_ms), int(units)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([int(units)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t[0]), int(u[0])


def gather_hashes(value: int, device=None, group=None):
    """All ranks' 63-bit trace hashes (e.g. of their eviction ids) on every rank — the host-side check that
    shards ran independent, deterministic work."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(value)]
    world = dist.get_world_size(group)
    out = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, torch.tensor([int(value) & ((1 << 63) - 1)], dtype=torch.int64, device=device), group=group)
    return [int(x[0]) for x in out]

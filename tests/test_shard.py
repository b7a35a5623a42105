"""The N>1 path on CPU: world_size-2 gloo process group, sequences sharded with no data-path collective;
only timings / counts / trace hashes are reduced (easykv_b200/shard.py, used by bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from easykv_b200 import shard


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_reduce_job_without_group_is_identity():
    assert shard.reduce_job(3.5, 10) == (3.5, 10)
    assert shard.gather_hashes(42) == [42]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(13, world, rank)
        # each rank "processes" its own sequences: the per-sequence result depends on the sequence id only
        local = [(s * 2654435761) % 1000003 for s in range(lo, hi)]
        ms, units = shard.reduce_job(10.0 + rank, hi - lo)
        hashes = shard.gather_hashes(sum(local))
        q.put((rank, lo, hi, ms, units, hashes))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ms0, u0, h0), (r1, lo1, hi1, ms1, u1, h1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 7, 7, 13)
    assert ms0 == ms1 == 11.0                      # max over ranks
    assert u0 == u1 == 13                          # sum over ranks
    assert h0 == h1 and len(h0) == 2
    assert sum(h0) == sum((s * 2654435761) % 1000003 for s in range(13))     # shards are disjoint and complete

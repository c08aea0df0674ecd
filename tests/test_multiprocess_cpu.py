"""world_size-2 gloo test (CPU) of the N > 1 plumbing: batch sharding, one-time key broadcast, max-over-ranks
timing and whole-job throughput."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lattigo_b200 import dist as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        key = torch.arange(1000, dtype=torch.int64) * 7 if rank == 0 else torch.zeros(1000, dtype=torch.int64)
        D.broadcast_key(key, src=0)
        lo, hi = D.shard_range(13, world, rank)
        t = D.max_over_ranks(0.5 + rank)              # rank 1 is slower
        thr = D.job_throughput(hi - lo, 0.5 + rank)
        out.put((rank, bool(torch.equal(key, torch.arange(1000, dtype=torch.int64) * 7)), lo, hi, t, thr))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                               # both ranks hold rank 0's key
    assert [(r[2], r[3]) for r in res] == [(0, 7), (7, 13)]     # contiguous balanced shards covering the batch
    assert all(abs(r[4] - 1.5) < 1e-9 for r in res)             # step time = slowest rank
    assert all(abs(r[5] - 13 / 1.5) < 1e-9 for r in res)        # whole-job throughput


def test_shard_range_properties():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1

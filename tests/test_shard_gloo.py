"""CPU (gloo, world_size 2): the N>1 host logic — batch sharding, the single weight-blob broadcast and the
throughput bookkeeping bench.py uses.  No GPU and no compute calls."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_batch_exactly():
    from hobot_stereonet_b200.shard import shard_range
    for n, world in [(32, 8), (64, 8), (64, 4), (10, 4), (1, 8), (0, 2), (7, 1)]:
        spans = [shard_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # contiguous, no gaps/overlap
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [shard_range(32, 8, r) for r in (0, 7)] == [(0, 4), (28, 32)]     # config 4: 4 pairs per GPU
    assert shard_range(10, 4, 1) == (3, 6)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hobot_stereonet_b200 import capi
        from hobot_stereonet_b200.shard import broadcast_blob, gather_counts, shard_range
        blob = capi.synthesize_weights(3, 1234) if rank == 0 else None       # only rank 0 "loads model_file"
        got = broadcast_blob(blob, src=0)
        lo, hi = shard_range(5, world, rank)
        total, ms = gather_counts(hi - lo, 10.0 + rank)
        q.put((rank, len(got), int(np.frombuffer(got, np.uint8).astype(np.uint64).sum()), got[:8], (lo, hi), total, ms))
    finally:
        dist.destroy_process_group()


def test_weight_broadcast_and_counts_world2(built_lib):
    from hobot_stereonet_b200 import capi
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = capi.synthesize_weights(3, 1234)
    for rank, n, csum, magic, span, total, ms in res:
        assert n == len(ref) and csum == int(np.frombuffer(ref, np.uint8).astype(np.uint64).sum())
        assert magic == b"SNB2WGT1"
        assert total == 5 and ms == 11.0                                     # sum of shards, max over ranks
    assert [r[4] for r in res] == [(0, 3), (3, 5)]

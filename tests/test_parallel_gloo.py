"""CPU, world_size 2, gloo: host-side logic of the N>1 path (sharding, the single weight-blob
broadcast, max-over-ranks timing).  The device side of the same code path runs under NCCL on
the box (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interactive_deep_colorization_b200.parallel import broadcast_blob, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 16, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 owns the packed weights; everyone else receives them with ONE broadcast
        blob = torch.arange(4096, dtype=torch.uint8) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
        broadcast_blob(blob, src=0)
        ok_blob = bool(torch.equal(blob, torch.arange(4096, dtype=torch.uint8)))
        # image sharding: every rank computes a checksum of its slice; the union must cover the batch
        start, count = shard_range(13, world, rank)
        data = np.arange(13, dtype=np.float64) ** 2
        part = torch.tensor([data[start:start + count].sum()], dtype=torch.float64)
        dist.all_reduce(part)                                   # test-only collective
        ok_shard = abs(part.item() - data.sum()) < 1e-9
        # step time = slowest rank
        t = max_over_ranks(10.0 + rank)
        q.put((rank, ok_blob, ok_shard, t))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_blob, ok_shard, t in res:
        assert ok_blob and ok_shard and t == 11.0

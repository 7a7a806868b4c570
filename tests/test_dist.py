"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharding + gather plumbing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_verde_b200 import dist as mdist


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 65536, 1000003):
        for G in (1, 2, 4, 8):
            r = [mdist.shard_range(B, k, G) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = mdist.shard_range(B, rank, world)
    full = torch.arange(B * 3, dtype=torch.float64).reshape(B, 3)
    mine = full[lo:hi] * 2.0                                  # "solve" of my shard
    got = mdist.gather_rows(mine)
    status = torch.zeros(hi - lo, dtype=torch.int32)
    iters = torch.full((hi - lo,), 10 + rank, dtype=torch.int32)
    stats = mdist.reduce_stats(status, iters)
    if rank == 0:
        torch.save({"got": got, "stats": stats}, out)
    dist.destroy_process_group()


def test_gather_and_stats_world_size_2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    B = 11
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
    r = torch.load(out)
    full = torch.arange(B * 3, dtype=torch.float64).reshape(B, 3) * 2.0
    assert torch.equal(r["got"], full)                       # bit-for-bit the single-process result
    assert r["stats"]["problems"] == B and r["stats"]["succeeded"] == B
    assert r["stats"]["iters_sum"] == 5 * 10 + 6 * 11 and r["stats"]["iters_max"] == 11

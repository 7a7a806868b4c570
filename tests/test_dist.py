"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharding + gather plumbing."""
import os
import socket

import numpy as np
import pytest
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
    # the packed one-collective form of the benchmark (equal shards): same bits, same statistics
    eq = full[rank * 4:(rank + 1) * 4] * 2.0
    pk = mdist.gather_packed([eq, eq[:, 0]], torch.zeros(4, dtype=torch.int32), torch.full((4,), 10 + rank, dtype=torch.int32))
    if rank == 0:
        torch.save({"got": got, "stats": stats, "pk_x": pk.outs[0].clone(), "pk_f": pk.outs[1].clone(), "pk_stats": pk.stats(),
                    "pk_iters": pk.iters.clone()}, out)
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
    assert torch.equal(r["pk_x"], full[:8]) and torch.equal(r["pk_f"], full[:8, 0])
    assert r["pk_stats"] == {"problems": 8, "succeeded": 8, "iters_sum": 4 * 10 + 4 * 11, "iters_max": 11}
    assert r["pk_iters"].tolist() == [10.0] * 4 + [11.0] * 4


# ---- T11 on real GPUs: NCCL gather of SOLVER output == the shards solved alone, bit for bit ----------------------
def _gpu_worker(rank, world, port, B, out):
    import mpc_verde_b200 as mv
    from mpc_verde_b200 import problems
    from tests import common
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    prob = problems.unicycle_multiple_shooting()
    solver = mv.nlpsol("solver", "ipopt", prob, {"ipopt": {"max_iter": 2000, "print_level": 0}})
    sp = solver.spec
    x0s, p = common.unicycle_batch(B)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    w0 = problems.cold_start(sp, x0s)
    lo, hi = mdist.shard_range(B, rank, world)
    sol = solver(x0=torch.as_tensor(w0[lo:hi]).to(dev), lbx=lbx, ubx=ubx, p=torch.as_tensor(p[lo:hi]).to(dev), outputs=("x", "f"))
    st, it = solver._last
    xs, fs = mdist.gather_rows(sol["x"]), mdist.gather_rows(sol["f"])
    stats = mdist.reduce_stats(st, it)
    pk = mdist.gather_packed([sol["x"], sol["f"]], st, it)        # bench.py's one-collective form (equal shards)
    if rank == 0:
        # every shard once more, alone on this GPU: the multi-GPU plumbing must not change one bit
        alone = []
        for r in range(world):
            a, b = mdist.shard_range(B, r, world)
            s1 = solver(x0=torch.as_tensor(w0[a:b]).to(dev), lbx=lbx, ubx=ubx, p=torch.as_tensor(p[a:b]).to(dev), outputs=("x", "f"))
            alone.append((s1["x"].clone(), s1["f"].clone(), solver._last[1].clone()))
        torch.save({"xs": xs.cpu(), "fs": fs.cpu(), "stats": stats, "pk_x": pk.outs[0].cpu(), "pk_f": pk.outs[1].cpu(),
                    "pk_stats": pk.stats(),
                    "x1": torch.cat([a[0] for a in alone]).cpu(), "f1": torch.cat([a[1] for a in alone]).cpu(),
                    "iters1": int(sum(int(a[2].sum()) for a in alone))}, out)
    dist.destroy_process_group()


@pytest.mark.gpu
def test_multi_gpu_gather_of_solver_output_is_bit_identical(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    B = 20000
    out = str(tmp_path / "r0.pt")
    mp.spawn(_gpu_worker, args=(2, port, B, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["xs"].shape == (B, 53) and torch.equal(r["xs"], r["x1"]) and torch.equal(r["fs"], r["f1"])
    assert r["stats"]["problems"] == B and r["stats"]["succeeded"] == B and r["stats"]["iters_sum"] == r["iters1"]
    assert torch.equal(r["pk_x"], r["x1"]) and torch.equal(r["pk_f"], r["f1"]) and r["pk_stats"] == r["stats"]

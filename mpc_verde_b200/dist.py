"""Multi-GPU plumbing: the batch of independent MPC problems shards by problem index, one
process per GPU, no collective on the hot path.  torch.distributed (NCCL over NVLink on the
GPU box, gloo in CPU tests) only gathers per-rank results and reduces statistics.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous block of problem indices owned by `rank`: [r*B/G, (r+1)*B/G)."""
    return (B * rank) // world, (B * (rank + 1)) // world


def gather_rows(t, group=None):
    """All-gather a [b_r, ...] tensor along dim 0 (ranks may own different b_r)."""
    if not (dist.is_available() and dist.is_initialized()):
        return t
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def reduce_stats(status, iters, group=None):
    """Small all-reduced statistics vector: #problems, #succeeded, sum and max of IPM iterations."""
    dev = status.device
    v = torch.tensor([status.numel(), int((status == 0).sum()), int(iters.sum())], dtype=torch.int64, device=dev)
    mx = torch.tensor([int(iters.max()) if iters.numel() else 0], dtype=torch.int64, device=dev)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return {"problems": int(v[0]), "succeeded": int(v[1]), "iters_sum": int(v[2]), "iters_max": int(mx[0])}


def gather_rows_equal(t, group=None):
    """All-gather when every rank owns the same number of rows (weak scaling): one collective, no host
    synchronisation, so it can run on a side stream underneath the next solve."""
    if not (dist.is_available() and dist.is_initialized()):
        return t
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def reduce_stats_device(status, iters, group=None):
    """reduce_stats without host round trips: returns device tensors [problems, succeeded, iters_sum], [iters_max]."""
    v = torch.stack([torch.tensor(status.numel(), dtype=torch.int64, device=status.device),
                     (status == 0).sum().to(torch.int64), iters.sum().to(torch.int64)])
    mx = iters.max().to(torch.int64).reshape(1) if iters.numel() else torch.zeros(1, dtype=torch.int64, device=status.device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return v, mx

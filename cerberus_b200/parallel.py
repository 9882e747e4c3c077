"""Batch split of independent windows across ranks.

Windows are independent, so the only multi-GPU pattern of this path is a contiguous split of the batch
index over ranks with no collective inside a solve; torch.distributed is used for rendezvous, barriers
and for gathering reports / timings (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of n_total windows for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local_rows, dist, device=None):
    """All-gather a [n_local, k] float64 array with possibly different n_local per rank -> [n_total, k] on every rank."""
    import torch
    world = dist.get_world_size()
    t = torch.as_tensor(np.ascontiguousarray(local_rows), dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx, t.shape[1]), dtype=torch.float64, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)


def max_over_ranks(value, dist, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

"""Multi-GPU host logic: the batch shards trivially (independent instances); no collective on
the solve path.  The only exchange is the closed-loop / sweep statistics (`all_reduce` of O(10) numbers)."""
from __future__ import annotations

import numpy as np


def shard_range(B, rank, world):
    """contiguous split of the batch dimension: rank r owns [lo, hi)"""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_stats(status, iters, flags, nAS):
    """fixed-size statistics vector of one shard: status histogram (5), flag histogram (4), sum iter, max iter, sum nAS, count"""
    status = np.asarray(status); iters = np.asarray(iters); flags = np.asarray(flags); nAS = np.asarray(nAS)
    v = np.zeros(13, dtype=np.float64)
    v[0:5] = np.bincount(status.clip(0, 4), minlength=5)[:5]
    v[5:9] = np.bincount(flags.clip(0, 3), minlength=4)[:4]
    v[9] = iters.sum()
    v[10] = iters.max() if iters.size else 0
    v[11] = nAS.sum()
    v[12] = status.size
    return v


def reduce_stats(v, dist=None):
    """sum-reduce everything except the max entry; `dist` = torch.distributed (NCCL on GPUs, gloo in the CPU tests)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return v
    import torch
    t = torch.as_tensor(v).clone()
    mx = t[10:11].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[10] = mx[0]
    return t.numpy() if not t.is_cuda else t.cpu().numpy()

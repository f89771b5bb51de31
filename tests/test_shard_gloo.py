"""world_size-2 gloo test of the multi-GPU host logic (sharding + statistics reduction), CPU only."""
import os

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, B, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tunempc_b200.sharding import local_stats, reduce_stats, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    status = rng.integers(0, 3, B); iters = rng.integers(1, 20, B); flags = rng.integers(0, 4, B); nAS = rng.integers(0, 18, B)
    lo, hi = shard_range(B, rank, world)
    v = reduce_stats(local_stats(status[lo:hi], iters[lo:hi], flags[lo:hi], nAS[lo:hi]), dist)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([v, [lo, hi]]))
    dist.destroy_process_group()


def test_shard_and_reduce_world2(tmp_path):
    from tunempc_b200.sharding import local_stats, shard_range
    B, world = 1001, 2
    assert [shard_range(B, r, world) for r in range(world)] == [(0, 501), (501, 1001)]
    assert [shard_range(7, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 7)]
    mp.spawn(_worker, args=(world, 29533, B, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    status = rng.integers(0, 3, B); iters = rng.integers(1, 20, B); flags = rng.integers(0, 4, B); nAS = rng.integers(0, 18, B)
    ref = local_stats(status, iters, flags, nAS)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert np.array_equal(got[:13], ref)          # reduction over shards == single-process statistics

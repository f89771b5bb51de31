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


def _worker_cl(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tunempc_b200.closed_loop_tools import reduce_rollout_stats, rollout_stats
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = _fake_log(rank)
    v = reduce_rollout_stats(rollout_stats(log, "c", l_ref=np.arange(4.0)), dist)
    np.save(os.path.join(out_dir, "cl%d.npy" % rank), v.numpy())
    dist.destroy_process_group()


def _fake_log(rank):
    import torch
    g = torch.Generator().manual_seed(rank)
    B, N = 5 + rank, 4
    return {"l": {"c": [torch.rand(B, dtype=torch.float64, generator=g) for _ in range(N)]},
            "h": {"c": [torch.rand((B, 3), dtype=torch.float64, generator=g) - 0.1 * (rank + 1) for _ in range(N)]},
            "status": {"c": [(torch.rand(B, generator=g) < 0.2).to(torch.int32) for _ in range(N)]}}


def test_rollout_stats_reduce_world2(tmp_path):
    """closed-loop statistics: the only collective of a sharded rollout (north_star: 'NCCL gather only for the
    closed-loop statistics'); gloo stands in for NCCL on CPU."""
    import torch
    from tunempc_b200.closed_loop_tools import rollout_stats
    world = 2
    mp.spawn(_worker_cl, args=(world, 29541, str(tmp_path)), nprocs=world, join=True)
    loc = [rollout_stats(_fake_log(r), "c", l_ref=np.arange(4.0)) for r in range(world)]
    ref = loc[0] + loc[1]
    ref[1] = max(loc[0][1], loc[1][1])
    ref[5] = max(loc[0][5], loc[1][5])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "cl%d.npy" % r))
        assert np.allclose(got, ref.numpy(), rtol=1e-14)
    assert ref[0] == 11 and ref[1] == 4 and ref[5] > 0

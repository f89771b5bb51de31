#!/usr/bin/env python
"""Closed-loop throughput of BASELINE.json config #4 (examples/unicycle): B = 2^16 rollouts x 100 MPC steps per GPU,
periodic tuned tracking MPC (p = N = 30), plant = model, through tunempc_b200.closed_loop_tools.closed_loop_sim.
Not the driver's bench line (bench.py measures the metric on the CSTR config); prints one JSON line for profiles/.
  python tools/bench_closed_loop.py [--batch 65536] [--steps 100]        (torchrun for N > 1: NCCL reduces the statistics)"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1 << 16)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tunempc_b200 import closed_loop_tools as clt
    from tunempc_b200.pmpc import Pmpc
    from tunempc_b200.problem import MpcProblem
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pb = MpcProblem.load(os.path.join(ROOT, "tests", "golden", "problem_unicycle.npz"))
    ctrl = Pmpc(pb, device=local)
    rng = np.random.default_rng(100 + rank)
    B = args.batch
    X0 = pb.wref[0, :pb.nx] + np.array([0.5, 0.1, 0.0, 0.0]) * rng.uniform(-1, 1, (B, pb.nx))   # examples/unicycle/main.py:172
    l_ref = np.array([pb.wref[k % pb.p, 4] ** 2 + pb.wref[k % pb.p, 0] ** 2 + 5 * pb.wref[k % pb.p, 1] ** 2 for k in range(args.steps)])
    clt.closed_loop_sim({"TUNEMPC": ctrl}, None, None, None, X0[:1024], 3)      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    log = clt.closed_loop_sim({"TUNEMPC": ctrl}, None, None, None, X0, args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    v = clt.rollout_stats(log, "TUNEMPC", l_ref=l_ref)
    it = torch.stack([t.to(torch.float64).mean() for t in ctrl.log["iter"]]).mean().reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        v = clt.reduce_rollout_stats(v, dist)                                   # the only data collective
        dist.all_reduce(it, op=dist.ReduceOp.SUM); it /= world
    if rank == 0:
        xerr = float((log["x"]["TUNEMPC"][-1] - torch.as_tensor(pb.wref[args.steps % pb.p, :pb.nx], device=dev)).abs().max())
        print(json.dumps({"metric": "tuned-MPC solves/sec (closed loop, fp64)", "value": world * B * args.steps / (float(ms) * 1e-3),
                          "unit": "solves/s", "n_gpus": world, "ms_total": float(ms),
                          "config": {"workload": "unicycle p=N=30 periodic tuned MPC, %d rollouts x %d steps per GPU, plant = model" % (B, args.steps)},
                          "stats": {"rollouts": float(v[0]), "steps": float(v[1]), "sum_stage_cost": float(v[2]),
                                    "transient_cost_mean_per_rollout": float(v[3] / v[0]), "non_converged_solves": float(v[4]),
                                    "sqp_iter_mean": float(it), "max_abs_tracking_error_at_end_rank0": xerr}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

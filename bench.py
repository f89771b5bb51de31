#!/usr/bin/env python
"""bench.py -- tuned-MPC solves/sec (batched Pmpc.step, fp64) on the CSTR config of BASELINE.json.

  python bench.py --gpus N --steps K --warmup W            B200 arm (under torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the oracle port of the reference loop on host cores

A "step" = one pass of the hot path over one batch: `ctrl.reset(); ctrl.step(X0)` for B = 2^20 seeded initial states
per GPU (the alpha-sweep loop of tunempc/closed_loop_tools.py:43-68 as one batched call).  Weak scaling: every rank
solves its own 2^20-instance shard, no collective on the solve path; NCCL is used only for the statistics.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # CPU arm: one single-threaded worker process per core -- keep the BLAS / OpenMP pools from oversubscribing the box
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = "1"

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tuned-MPC solves/sec (batched step, fp64)"
UNIT = "solves/s"


def sample_x0(pb, B, seed):
    """cA-direction sweep alpha in [-0.1, 1.0] of examples/cstr/main.py:124-131 plus
    1e-2*|x_s| jitter on the other states (same generator as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    xs = pb.wref[0, :pb.nx]
    alpha = rng.uniform(-0.1, 1.0, B)
    X0 = np.tile(xs, (B, 1))
    X0[:, 0] += alpha * (1.0 - xs[0])
    X0[:, 1:] += 1e-2 * np.abs(xs[1:]) * rng.uniform(-1, 1, (B, pb.nx - 1))
    return X0


def load_problem():
    from tunempc_b200.problem import MpcProblem
    return MpcProblem.load(os.path.join(ROOT, "tests", "golden", "problem_cstr.npz"))


def algorithmic_flops_per_stage_lin(exact=True):
    """fp64 add/mul/div = 1, FMA = 2, op counts of the generated model code (modelgen)."""
    from tunempc_b200 import configs, modelgen
    import tempfile
    m = configs.cstr()["model"]
    with tempfile.TemporaryDirectory() as d:
        oc = modelgen.generate_header(m, os.path.join(d, "m.h"))
    nx, nu, M = m.nx, m.nu, m.rk_steps
    nz = nx + nu
    c_f, c_J = oc["c_f"], oc["c_J"] - oc["c_f"]
    c_H = (oc["c_H"] - oc["c_J"]) + oc["c_bilin"]
    f_gn = M * (4 * (c_f + c_J + 2 * nx * nx * nz) + 16 * nx * (1 + nz))
    f_ex = f_gn + M * 4 * (c_H + 2 * nz * nz * nx + 2 * nx * nz * nz)
    f_dyn = M * (4 * c_f + 16 * nx)
    return (f_ex if exact else f_gn), f_dyn


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.stop_ev = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ------------------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    idx_list, seed = args
    from threadpoolctl import threadpool_limits
    from oracle import reference_port as rp
    pb = load_problem()
    ctrl = rp.Pmpc(pb)
    X0 = sample_x0(pb, max(idx_list) + 1, seed)
    t = time.perf_counter()
    with threadpool_limits(limits=1):
        for i in idx_list:
            ctrl.reset()                                   # closed_loop_tools.py:68
            ctrl.step(X0[i])                               # closed_loop_tools.py:56
    return time.perf_counter() - t, len(idx_list)


def run_reference_arm(args):
    """CPU arm: the oracle port of `for x0: ctrl.reset(); ctrl.step(x0)` on all host cores (one process per core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import reference_port as rp
    rp.build(("cstr",))
    cores = len(os.sched_getaffinity(0))
    per_core = 6                                   # ~2 s of CPU work per core and step
    n = cores * per_core
    chunks = [(list(range(c, n, cores)), 1000) for c in range(cores)]
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_oracle_worker, [(c[0], 1000 + s) for c in chunks])
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    tot = sum(times)
    val = n * len(times) / tot
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "cstr N=20 tuned NMPC (examples/cstr), exact Hessian, reset+step per x0",
                       "sample": "%d x0 per step (bounded sample of the B=2^20-per-GPU workload of the B200 arm)" % n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d instances per step (oracle port: numpy + C stage functions + qpOASES_e), %d steps" % (n, len(times))},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=1 << 20, help="instances per GPU per step")
    ap.add_argument("--hessian", default="exact")
    ap.add_argument("--cpu-sample", type=int, default=40, help="x0 of the batch timed on the CPU oracle port (about 12 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from tunempc_b200.pmpc import Pmpc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pb = load_problem()
    pb.hessian_approximation = args.hessian
    ctrl = Pmpc(pb, device=local)
    B = args.batch
    W = max(args.warmup, 3)
    K = args.steps
    nbatches = 2
    X0_host = [torch.from_numpy(sample_x0(pb, B, 100 + 17 * rank + i)).pin_memory() for i in range(nbatches)]
    X0_dev = [x.to(dev) for x in X0_host]
    U_host = torch.empty((B, pb.nu), dtype=torch.float64).pin_memory()
    st_host = torch.empty(B, dtype=torch.int32).pin_memory()
    peak_tf = ctrl.fp64_peak_tflops()

    # ---- device-resident arm: inputs already in HBM --------------------------------------------------------
    def dev_step(i):
        ctrl.reset()
        return ctrl.step(X0_dev[i % nbatches], outputs="u0")

    for i in range(W):
        dev_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lin_ms = qp_ms = post_ms = 0.0
    n_lin = n_qp = n_it = n_launch = n_dyn = 0
    for i in range(K):
        dev_step(i)
        t = ctrl.timing()
        c = ctrl.counters()
        lin_ms += t["lin_ms"]; qp_ms += t["qp_ms"]; post_ms += t["post_ms"]
        n_lin += c["stage_linearisations"]; n_qp += c["qp_solves"]; n_it += c["sqp_iterations"]
        n_launch += c["kernel_launches"]; n_dyn += c["ls_dynamics_evals"]
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    status = ctrl.status
    flags = ctrl.log["flags"][-1]
    stat_hist = torch.bincount(status.to(torch.int64), minlength=5)[:5].to(torch.float64)
    # per-bit counts: [no flag, GN fallback (1), damped step (2), GN re-solve (4), non-convex primal step (8)]
    fl_hist = torch.stack([(flags == 0).sum()] + [((flags & b) != 0).sum() for b in (1, 2, 4, 8)]).to(torch.float64)

    # ---- end-to-end arm: host buffers through the C ABI (tmpc_step_host), H2D + D2H inside the timed region ----
    x_np = [x.numpy() for x in X0_host]

    def host_step(i):
        ctrl.reset()
        u = ctrl.step(x_np[i % nbatches], outputs="u0")
        return u

    host_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_step(i)
    torch.cuda.synchronize()
    ms_e2e_local = 1e3 * (time.perf_counter() - t0)
    barrier()
    sampler.stop_ev.set()
    sampler.join(timeout=2)

    tvec = torch.tensor([ms_dev, ms_e2e_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)              # max over ranks
        dist.all_reduce(stat_hist, op=dist.ReduceOp.SUM)         # closed-loop statistics: the only collectives
        dist.all_reduce(fl_hist, op=dist.ReduceOp.SUM)
    ms_dev, ms_e2e = float(tvec[0]), float(tvec[1])

    if rank == 0:
        f_lin, f_dyn = algorithmic_flops_per_stage_lin(args.hessian == "exact")
        ach = (n_lin * f_lin) / (lin_ms * 1e-3) / 1e12 if lin_ms > 0 else 0.0
        # DRAM traffic of the dominant kernels from the committed ncu --set full captures (profiles/r01e_summary.md):
        # k_lin2 440.0 MB read + 972.8 MB written per launch of 131072 x 20 stage tasks; k_qp_thread 30.3 + 6.2 GB per
        # launch of 131072 QPs.  Algorithmic bytes: a stage task reads (x,u,lam_dyn) and writes its 49-double record;
        # a QP reads its N records + w and writes (d, lam).
        lin_traffic_per_task = (439.988992e6 + 972.843008e6) / (131072 * 20)
        lin_alg_bytes_per_task = 8.0 * (pb.nx + pb.nu + pb.nx) + 8.0 * (pb.nx + pb.nx * pb.nz + pb.nz * (pb.nz + 1) // 2)
        qp_traffic_per_qp = (30.312675e9 + 6.179678e9) / 131072
        qp_alg_bytes = 8.0 * (pb.N * (pb.nx + pb.nx * pb.nz + pb.nz * (pb.nz + 1) // 2) + 2 * pb.n_w + pb.n_g + pb.nx)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak = float(json.load(fh)["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            hbm_peak, hbm_src = 6550.0, "fallback (B200_PROFILING.md)"
        qp_ach = n_qp * qp_alg_bytes / (qp_ms * 1e-3) / 1e9 if qp_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": world * B * K / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cstr N=20 tuned NMPC (examples/cstr), %s Hessian, B=%d x0 per GPU per step, reset+step"
                                   % (args.hessian, B),
                       "l2": "working set %.1f GB per step >> L2; %d alternating input batches" % (B * 14.5e3 / 1e9, nbatches),
                       "x0": "cA sweep alpha~U(-0.1,1.0) + 1e-2 jitter, seed 100+17*rank+i"},
            "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * pb.nx * 8,
                    "d2h_bytes_per_step": B * pb.nu * 8 + 3 * B * 4},
            "gpu_launches": int(n_launch),
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "kernel": "k_lin2", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": ach / peak_tf if peak_tf else None,
                         "traffic": lin_traffic_per_task * 32 * ((B * pb.N + 31) // 32),
                         "traffic_note": "ncu dram read+write per full-batch launch (%.0f B per stage task measured at B=131072, "
                                         "algorithmic %.0f B: DRAM traffic = the records)" % (lin_traffic_per_task, lin_alg_bytes_per_task),
                         "peak_source": "in-run DFMA micro-benchmark (tmpc_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                         "flops_per_stage_linearisation": f_lin, "stage_linearisations": int(n_lin),
                         "kernel_ms": {"k_lin": lin_ms, "k_qp": qp_ms, "k_post": post_ms, "step_total": ms_dev},
                         "hbm_GBps_boundary_io": (B * K * 8 * (pb.nx + pb.nu) + 0.0) / (ms_dev * 1e-3) / 1e9},
            "roofline_qp": {"bound": "hbm", "kernel": "k_qp_thread (+ k_qp0, k_qp)", "achieved": qp_ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": qp_ach / hbm_peak, "traffic": qp_traffic_per_qp * B,
                            "note": "algorithmic %.0f B per QP (records + w in, d + lam out); measured DRAM traffic %.0f B per QP: the "
                                    "lane-interleaved Riccati / working-set workspace streams through HBM" % (qp_alg_bytes, qp_traffic_per_qp),
                            "peak_source": hbm_src},
            "stats": {"sqp_iter_mean": n_it / (B * K), "qp_solves": int(n_qp), "ls_dynamics_evals": int(n_dyn),
                      "status_hist": [int(v) for v in stat_hist.tolist()], "flags_hist": [int(v) for v in fl_hist.tolist()],
                      "flags_hist_keys": ["none", "gn_fallback", "damped", "gn_resolve", "nonconvex_step"]},
        }
        # ---- CPU baseline: the oracle port on a bounded sample of the same workload, 1 core ----
        try:
            from oracle import reference_port as rp
            oc = rp.Pmpc(load_problem())
            n = args.cpu_sample
            xs = x_np[0][:n]
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=1):              # "cores": 1 means one thread, BLAS included
                t0 = time.perf_counter()
                for i in range(n):
                    oc.reset()
                    oc.step(xs[i])
                dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "first %d x0 of the same batch, oracle port (numpy + C stage functions + qpOASES_e)" % n}
        except Exception as e:   # the oracle is a checker, its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

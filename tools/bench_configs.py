#!/usr/bin/env python
"""Throughput of one reset + step at the batch sizes BASELINE.json names for the other configs (parity-test cases of the
bench; not the driver's line): evaporation 2^18, unicycle 2^16 (single step; the closed loop is tools/bench_closed_loop.py),
dims9 (AWE dimensions) 2^14, economic CSTR 2^18, lq 2^10.  One JSON line per config."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from tunempc_b200.pmpc import Pmpc
from tunempc_b200.problem import MpcProblem
from make_golden import sample_x0

for fixture, name, B in (("lq", "lq", 1 << 10), ("evaporation", "evaporation", 1 << 18), ("unicycle", "unicycle", 1 << 16),
                         ("dims9", "dims9", 1 << 14), ("cstr_economic", "cstr", 1 << 18), ("chain", "chain", 1 << 16)):
    pb = MpcProblem.load(os.path.join(ROOT, "tests", "golden", "problem_%s.npz" % fixture))
    ctrl = Pmpc(pb, device=0)
    X0 = torch.tensor(sample_x0(name, pb, B, seed=11), device="cuda:0")
    for _ in range(2):
        ctrl.reset(); ctrl.step(X0, outputs="u0")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 3
    for _ in range(K):
        ctrl.reset(); ctrl.step(X0, outputs="u0")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    st = np.bincount(ctrl.status.cpu().numpy(), minlength=5)[:5].tolist()
    t = ctrl.timing()
    print(json.dumps({"config": fixture, "mpc_type": pb.mpc_type, "B": B, "ms_per_step": ms, "solves_per_s": B / (ms * 1e-3),
                      "sqp_iter_mean": float(ctrl.log["iter"][-1].float().mean()), "status_hist": st,
                      "kernel_ms": {k: round(v, 2) for k, v in t.items()}}))

#!/usr/bin/env python
"""Parity sweep at a size the committed fixtures do not reach: B random benchmark instances solved by the CUDA path
(shared-table first QP and generic route), by the CPU twin of the kernels (same algorithm, host compiler) and -- on a random
subset -- by the oracle.  Prints one JSON summary (profiles/)."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import load_problem, sample_x0
from tunempc_b200.pmpc import Pmpc
from tunempc_b200.problem import build_tables
from twin.twin import Twin
from oracle import reference_port as rp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_oracle = 64
pb = load_problem()
X0 = sample_x0(pb, B, 2024)
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1.0)
out = {"config": "cstr N=20 tuned NMPC, exact Hessian", "B": B}
res = {}
for tag, env in (("shared_first_qp", {"TMPC_QP0_MIN": "2"}), ("generic", {"TMPC_QP0_MIN": "-1"})):
    os.environ.update(env)
    c = Pmpc(load_problem(), device=0)
    U = c.step(torch.tensor(X0, device="cuda:0")).cpu().numpy()
    res[tag] = dict(u=U, w=c.w_sol.cpu().numpy(), lam=c.lam_g.cpu().numpy(), st=c.status.cpu().numpy(),
                    it=c.log["iter"][-1].cpu().numpy(), fl=c.log["flags"][-1].cpu().numpy())
    del c
t = time.time()
tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
tw.reset(B)
o = tw.step(X0, shared_first_qp=True)
out["twin_seconds"] = time.time() - t
g = res["shared_first_qp"]
ok = (g["st"] == 0) & (o["status"] == 0)
out["gpu_vs_twin"] = {"status_equal": int((g["st"] == o["status"]).sum()), "iter_equal": int((g["it"] == o["iter"]).sum()),
                      "flags_equal": int((g["fl"] == o["flags"]).sum()),
                      "u0_max_rel": float(rel(g["u"][ok], o["u0"][ok]).max()), "w_max_rel": float(rel(g["w"][ok], o["w"][ok]).max()),
                      "active_sets_equal": int(sum(np.array_equal(g["lam"][b] != 0, o["lam"][b] != 0) for b in range(B)))}
h = res["generic"]
ok2 = (g["st"] == 0) & (h["st"] == 0)
out["shared_vs_generic_first_qp"] = {"iter_equal": int((g["it"] == h["it"]).sum()), "u0_max_rel": float(rel(g["u"][ok2], h["u"][ok2]).max()),
                                     "active_sets_equal": int(sum(np.array_equal(g["lam"][b] != 0, h["lam"][b] != 0) for b in range(B)))}
rng = np.random.default_rng(5)
sel = rng.choice(B, n_oracle, replace=False)
oc = rp.Pmpc(pb)
uerr, werr, as_eq, it_o = [], [], 0, []
t = time.time()
ineq = np.concatenate([np.arange(pb.g_h(k).start, pb.g_h(k).stop) for k in range(pb.N)])
for b in sel:
    oc.reset()
    uo = oc.step(X0[b])
    uerr.append(rel(g["u"][b], uo).max()); werr.append(rel(g["w"][b], oc.w_sol).max())
    as_eq += int(np.array_equal(g["lam"][b][ineq] != 0, oc.lam_g[ineq] != 0)); it_o.append(oc.log["iter"][-1])
out["gpu_vs_oracle"] = {"n": n_oracle, "u0_max_rel": float(max(uerr)), "w_max_rel": float(max(werr)), "active_sets_equal": as_eq,
                        "oracle_iter_mean": float(np.mean(it_o)), "gpu_iter_mean_same_instances": float(g["it"][sel].mean()),
                        "oracle_seconds": time.time() - t}
out["gpu_status_hist"] = np.bincount(g["st"], minlength=5)[:5].tolist()
print(json.dumps(out))

"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck): CSTR through the generic route, the
shared-table first QP (k_qp0*) and the thread-per-instance QP kernel, unicycle (periodic, 2 steps: shift), evaporation
(collocation), lq (discrete), economic controller, plant step and stage log."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tunempc_b200.pmpc import Pmpc
from tunempc_b200.problem import MpcProblem


def run(fixture, B, steps=1, **env):
    for k, v in env.items():
        os.environ[k] = v
    pb = MpcProblem.load(os.path.join(ROOT, "tests", "golden", "problem_%s.npz" % fixture))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % fixture))
    ctrl = Pmpc(pb, device=0)
    X = torch.tensor(gold["X0"][:B], device="cuda:0")
    for _ in range(steps):
        U = ctrl.step(X)
        l, hv = ctrl.stage_log(X, U)
        X = ctrl.plant_step(X, U)
    torch.cuda.synchronize()
    st = ctrl.status.cpu().numpy()
    print(fixture, env, "status", np.bincount(st), "u0[0]", U[0].cpu().numpy())
    for k in env:
        os.environ.pop(k)
    del ctrl


run("cstr", 40)
run("cstr", 40, TMPC_QP0_MIN="2")
run("cstr", 40, TMPC_QP_THREAD_MIN="1", TMPC_QP_MODE="t")
run("cstr", 24, TMPC_LIN_MODE="1")
run("unicycle", 8, steps=2)
run("evaporation", 8)
run("lq", 16)
run("cstr_economic", 8)
print("sanitize_small done")

"""tmpc_step_async / tmpc_wait (include/tmpc.h): the non-blocking form of the batched step.  Runs last (file name) and in a
child process with a timeout, so that a dead-locked worker thread fails this test instead of hanging the suite."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_problem
from tunempc_b200 import configs
from tunempc_b200.pmpc import Pmpc
pb = load_problem("cstr")
ctrl = Pmpc(pb, device=0)
X_host = torch.tensor(configs.sample_x0("cstr", pb, 8192, 7))
U_sync = ctrl.step(X_host.to("cuda:0")).clone()
it_sync = ctrl.log["iter"][-1].clone()
ctrl.reset()
torch.cuda.synchronize()
Xp = X_host.pin_memory()
X = torch.empty_like(X_host, device="cuda:0")
X.copy_(Xp, non_blocking=True)                   # producer of X0 on the stream, not synchronised
U = ctrl.step_async(X)
was_busy = ctrl.busy()
try:
    ctrl.step_async(X)
    second = "accepted"
except RuntimeError:
    second = "refused"                           # one step in flight per controller
Un = ctrl.wait()
assert Un is U and not ctrl.busy()
assert was_busy, "the call did not return before the SQP loop had finished"
assert second == "refused"
assert torch.equal(U, U_sync), "asynchronous result differs from the synchronous step"
assert torch.equal(ctrl.log["iter"][-1], it_sync) and bool((ctrl.status == 0).all()) and ctrl.index == 1
U2 = ctrl.step_async(ctrl.plant_step(X, U))      # closed loop
ctrl.wait()
assert bool((ctrl.status == 0).all()) and ctrl.index == 2 and bool(torch.isfinite(U2).all())
print("ASYNC_OK")
"""


@pytest.mark.gpu
def test_step_async_matches_synchronous_step(built):
    code = CHILD % (ROOT, os.path.join(ROOT, "tests"))
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=int(os.environ.get("TMPC_ASYNC_TEST_TIMEOUT", "240")))
    except subprocess.TimeoutExpired:
        pytest.fail("tmpc_step_async dead-locked (child process killed on timeout)")
    assert r.returncode == 0 and "ASYNC_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

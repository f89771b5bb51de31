#!/bin/bash
# A/B: consumer-role placement in k_lin2 (lightest role on the producer's scheduler), and step latency at small batches
run() { python bench.py --steps 2 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['kernel_ms'], round(d['roofline']['frac'],4))"; }
run new_roles
cp tunempc_b200/libtmpc_cstr.so /tmp/new.so; cp tunempc_b200/libtmpc_cstr_base.so tunempc_b200/libtmpc_cstr.so
run old_roles
cp /tmp/new.so tunempc_b200/libtmpc_cstr.so
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
from bench import load_problem, sample_x0
from tunempc_b200.pmpc import Pmpc
pb = load_problem()
for B in (1, 32, 1024, 32768):
    ctrl = Pmpc(pb, device=0)
    X0 = torch.tensor(sample_x0(pb, B, 3), device="cuda:0")
    for _ in range(3):
        ctrl.reset(); ctrl.step(X0, outputs="u0")
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        ctrl.reset(); ctrl.step(X0, outputs="u0")
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print("latency B=%d: %.2f ms per reset+step (%.0f solves/s), iter mean %.2f" % (B, dt * 1e3, B / dt, ctrl.log["iter"][-1].float().mean().item()))
PY

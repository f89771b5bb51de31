"""GPU: run one CSTR step at B = 2^20 and save the initial states of the instances that do not converge (status != 0)
or need more than 12 iterations -> gpurun_out/stragglers.npz (analysed offline with the CPU twin and the oracle)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_problem, sample_x0
from tunempc_b200.pmpc import Pmpc
pb = load_problem()
pb.max_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 60
ctrl = Pmpc(pb, device=0)
B = 1 << 20
X0 = sample_x0(pb, B, 100)
U = ctrl.step(torch.tensor(X0, device="cuda:0"), outputs="u0")
st = ctrl.status.cpu().numpy(); it = ctrl.log["iter"][-1].cpu().numpy(); fl = ctrl.log["flags"][-1].cpu().numpy()
sel = np.where((st != 0) | (it > 12))[0]
print("status hist", np.bincount(st), "iter hist", np.bincount(it), "selected", len(sel))
np.savez("gpurun_out/stragglers.npz", X0=X0[sel], status=st[sel], iter=it[sel], flags=fl[sel], idx=sel)

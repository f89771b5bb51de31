"""Tuner.create_mpc(..., opts={'slack_flag': 'active'}) end to end on the GPU (tunempc/tuner.py:171-177, preprocessing.py:120-155):
soft constraints through the user API, on the model-library variant evaporation_sc1, against the live oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tuner_soft_constraints_gpu(built):
    import torch
    from oracle import reference_port as rp
    from tunempc_b200 import configs
    from tunempc_b200.tuner import Tuner
    t = Tuner(configs.evaporation(), p=1)
    t.solve_ocp()
    t.convexify(rho=1e-3)
    ctrl = t.create_mpc("tuned", 30, opts={"slack_flag": "active"})
    pb = ctrl.problem
    assert pb.name == "evaporation_sc1" and pb.nsc == 1 and pb.nh == 6 and pb.scost[0] > 1e3
    xs = t.w_sol[0, :2]
    X0 = np.array([xs + [0.3, 0.5], [24.7, xs[1] - 0.4], [24.2, xs[1] + 0.2], xs + [0.05, -0.8]])
    U = ctrl.step(torch.tensor(X0, device="cuda:0")).cpu().numpy()
    assert (ctrl.status.cpu().numpy() == 0).all()
    w = ctrl.w_sol.cpu().numpy()
    assert np.isclose(w[1][pb.iusc(0)][0], 0.3, atol=1e-8) and np.isclose(w[2][pb.iusc(0)][0], 0.8, atol=1e-8)   # the slack absorbs x0 < bound
    oc = rp.Pmpc(pb)
    for b in range(4):
        oc.reset()
        uo = oc.step(X0[b])
        assert np.max(np.abs(U[b] - uo) / np.maximum(np.abs(uo), 1.0)) < 1e-6, b
        assert ctrl.log["iter"][-1].cpu().numpy()[b] == oc.log["iter"][-1]
    with pytest.raises(ValueError):
        t.create_mpc("tuned", 30, opts={"slack_flag": "some"})

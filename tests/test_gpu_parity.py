"""GPU parity tests (B200): the CUDA path, called through the C ABI (tunempc_b200.pmpc.Pmpc -> libtmpc_<model>.so),
against the oracle's golden outputs, a live oracle, and size-independent properties at large batch."""
import numpy as np
import pytest

from conftest import load_golden, load_problem

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))


@pytest.fixture(scope="module")
def torch_mod(built):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _ctrl(name, **kw):
    from tunempc_b200.pmpc import Pmpc
    pb = load_problem(name)
    for k, v in kw.items():
        setattr(pb, k, v)
    return Pmpc(pb, device=0), pb


def test_lq_golden(torch_mod):
    torch = torch_mod
    ctrl, pb = _ctrl("lq")
    gold = load_golden("lq")
    U = ctrl.step(torch.tensor(gold["X0"], device="cuda:0"))
    assert (ctrl.status.cpu().numpy() == 0).all()
    assert (ctrl.log["iter"][-1].cpu().numpy() == 1).all()                      # LQ: one iteration is exact
    # terminal rows are enforced through their multipliers (Schur complement); no penalty parameter limits the accuracy
    assert _relerr(U.cpu().numpy(), gold["u0_t6"]) < 1e-12
    assert _relerr(ctrl.w_sol.cpu().numpy(), gold["w_t6"]) < 2e-11
    assert _relerr(ctrl.lam_g.cpu().numpy(), gold["lam_t6"]) < 1e-10
    G = np.array([-0.08241103740895, -0.188345092908991, 0.225692606094775])    # SURVEY.md 8(c) known answer
    assert np.allclose(U.cpu().numpy()[:, 0], gold["X0"] @ G, atol=1e-9)


@pytest.mark.parametrize("tag,tol,q0min,qpmode", [("t6", 1e-6, None, None), ("t9", 1e-9, None, None), ("t6", 1e-6, 2, None),
                                                  ("t9", 1e-9, 2, None), ("t6", 1e-6, None, "t"), ("t9", 1e-9, 2, "t")])
def test_cstr_golden(torch_mod, tag, tol, q0min, qpmode, monkeypatch):
    """every production QP kernel meets the oracle: warp per instance (k_qp, default at this batch size), thread per
    instance (k_qp_thread, qpmode 't'), and the shared-table first QP (k_qp0, q0min = 2)"""
    torch = torch_mod
    if q0min is not None:
        monkeypatch.setenv("TMPC_QP0_MIN", str(q0min))   # first QP through the shared tables (k_qp0) even at B = 48
    if qpmode is not None:
        monkeypatch.setenv("TMPC_QP_MODE", qpmode)
        monkeypatch.setenv("TMPC_QP_THREAD_MIN", "1")
    ctrl, pb = _ctrl("cstr", tol=tol)
    gold = load_golden("cstr")
    U = ctrl.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    st = ctrl.status.cpu().numpy()
    assert (st == 0).all(), np.bincount(st)
    assert _relerr(U, gold["u0_" + tag]) < 1e-6                                  # north_star: 1e-6 relative on u0
    w = ctrl.w_sol.cpu().numpy()
    assert _relerr(w, gold["w_" + tag]) < (1e-5 if tag == "t6" else 1e-6)        # predicted trajectory
    lam = ctrl.lam_g.cpu().numpy()
    for b in range(lam.shape[0]):                                                # identical active sets
        assert set(np.nonzero(lam[b])[0]) == set(np.nonzero(gold["lam_" + tag][b])[0]), b
    assert np.array_equal(ctrl.log["nAS"][-1].cpu().numpy(), gold["nAS_" + tag])
    fl = ctrl.log["flags"][-1].cpu().numpy()
    clean = (fl & 13) == 0                                                       # convex QPs throughout: the oracle's iteration path
    assert clean.any()
    assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy()[clean], gold["iter_" + tag][clean])
    # g_sol: constraint values at the solution
    g = ctrl.g_sol.cpu().numpy()
    for k in range(pb.N):
        assert np.abs(g[:, pb.g_dyn(k)]).max() < 10 * tol and g[:, pb.g_h(k)].min() > -10 * tol
    assert np.abs(g[:, pb.g_init()]).max() < 10 * tol and np.abs(g[:, pb.g_term()]).max() < 10 * tol


def test_cstr_live_oracle_and_host_path(torch_mod):
    torch = torch_mod
    from oracle import reference_port as rp
    ctrl, pb = _ctrl("cstr")
    rng = np.random.default_rng(77)
    xs = pb.wref[0, :4]
    X0 = np.tile(xs, (6, 1))
    X0[:, 0] += rng.uniform(-0.1, 1.0, 6) * (1.0 - xs[0])
    U_host = ctrl.step(X0)                                                       # numpy in -> tmpc_step_host
    w_host = ctrl.w_sol.copy()
    ctrl.reset()
    U_dev = ctrl.step(torch.tensor(X0, device="cuda:0")).cpu().numpy()           # device tensors -> tmpc_step
    assert np.array_equal(U_host, U_dev) and np.array_equal(w_host, ctrl.w_sol.cpu().numpy())
    oc = rp.Pmpc(pb)
    for b in range(6):
        oc.reset()
        uo = oc.step(X0[b])
        assert _relerr(U_dev[b], uo) < 1e-6
        assert set(np.nonzero(ctrl.lam_g.cpu().numpy()[b])[0]) == set(np.nonzero(oc.lam_g)[0])
    # reference single-instance semantics: (nx,) in -> (nu,) out; reset-then-step idempotent (P5)
    ctrl.reset()
    u1 = ctrl.step(X0[0])
    ctrl.reset()
    u2 = ctrl.step(X0[0].reshape(4, 1))
    assert u1.shape == (2,) and u2.shape == (2, 1) and np.array_equal(u1, u2[:, 0]) and np.array_equal(u1, U_dev[0])


def test_cstr_reference_point(torch_mod):
    torch = torch_mod
    ctrl, pb = _ctrl("cstr")
    X0 = torch.tensor(np.tile(pb.wref[0, :4], (3, 1)), device="cuda:0")
    U = ctrl.step(X0).cpu().numpy()                                              # P1: step(x_ref) = u_ref
    assert np.allclose(U, pb.wref[0, 4:], rtol=1e-9)
    assert (ctrl.log["iter"][-1].cpu().numpy() == 1).all()                       # always >= 1 QP (sqp_method.py:145-146)


def test_cstr_closed_loop_and_shift(torch_mod):
    torch = torch_mod
    ctrl, pb = _ctrl("cstr")
    gold = load_golden("cstr")
    X = torch.tensor(gold["cl_X"][:, 0], device="cuda:0")
    for s in range(5):
        U = ctrl.step(X)
        assert _relerr(U.cpu().numpy(), gold["cl_U"][:, s]) < 1e-6, s
        X = ctrl.plant_step(X, U)                                                # closed_loop_tools.py:102
        assert _relerr(X.cpu().numpy(), gold["cl_X"][:, s + 1]) < 1e-6, s
    assert ctrl.index == 5


def test_gauss_newton_same_solution(torch_mod):
    torch = torch_mod
    a, pb = _ctrl("cstr", tol=1e-9)
    b, _ = _ctrl("cstr", tol=1e-9, hessian_approximation="gauss_newton")
    gold = load_golden("cstr")
    X0 = torch.tensor(gold["X0"][:16], device="cuda:0")
    Ua, Ub = a.step(X0).cpu().numpy(), b.step(X0).cpu().numpy()
    ok = (a.status.cpu().numpy() == 0) & (b.status.cpu().numpy() == 0)
    assert ok.sum() >= 12
    assert _relerr(Ua[ok], Ub[ok]) < 1e-6                                        # P3: Hessian-mode independence


def test_large_batch_properties(torch_mod):
    """size-independent properties at 2^16 instances: every instance converges, the result of an instance does not
    depend on its neighbours (bitwise), sharding the batch changes nothing (bitwise), KKT conditions hold."""
    torch = torch_mod
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import sample_x0
    ctrl, pb = _ctrl("cstr")
    B = 1 << 16
    X0 = torch.tensor(sample_x0(pb, B, 5), device="cuda:0")
    U = ctrl.step(X0)
    st = ctrl.status.cpu().numpy()
    assert (st == 0).mean() >= 0.99, np.bincount(st)
    w, lam, g = ctrl.w_sol, ctrl.lam_g, ctrl.g_sol
    ok = torch.tensor(st == 0, device="cuda:0")
    for k in range(pb.N):
        assert g[ok][:, pb.g_dyn(k)].abs().max().item() < 1e-5
        assert g[ok][:, pb.g_h(k)].min().item() > -1e-5
        lh = lam[ok][:, pb.g_h(k)]
        assert lh.max().item() <= 0.0                                            # multiplier sign
        assert (lh * g[ok][:, pb.g_h(k)]).abs().max().item() < 1e-4              # complementarity
    c2, _ = _ctrl("cstr")
    idx = torch.arange(1000, 1000 + 2048, device="cuda:0")
    U2 = c2.step(X0[idx].contiguous())
    assert torch.equal(U2, U[idx]) and torch.equal(c2.w_sol, w[idx])             # neighbours / sharding invariance
    # a batch below TMPC_QP0_MIN takes the per-instance route for its first QP instead of the shared tables:
    # same answers to round-off
    c3, _ = _ctrl("cstr")
    idx = torch.arange(5000, 5000 + 512, device="cuda:0")
    U3 = c3.step(X0[idx].contiguous()).cpu().numpy()
    ok3 = (c3.status.cpu().numpy() == 0) & (st[5000:5512] == 0)
    assert _relerr(U3[ok3], U[idx].cpu().numpy()[ok3]) < 1e-6


def test_unicycle_periodic_golden_and_closed_loop(torch_mod):
    """config #4 (examples/unicycle): p = N = 30 periodic reference with time-varying tuned H, projected terminal
    constraint; open-loop golden outputs, then the closed loop across the period wrap with the device plant step."""
    torch = torch_mod
    ctrl, pb = _ctrl("unicycle")
    gold = load_golden("unicycle")
    U = ctrl.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert (ctrl.status.cpu().numpy() == 0).all()
    assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["iter_t6"])
    assert _relerr(U, gold["u0_t6"]) < 1e-6
    assert _relerr(ctrl.w_sol.cpu().numpy(), gold["w_t6"]) < 1e-6
    assert _relerr(ctrl.lam_g.cpu().numpy(), gold["lam_t6"]) < 1e-5
    c9, _ = _ctrl("unicycle", tol=1e-9)
    U9 = c9.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert _relerr(U9, gold["u0_t9"]) < 1e-8 and _relerr(c9.w_sol.cpu().numpy(), gold["w_t9"]) < 1e-8
    ctrl.reset()
    X = torch.tensor(gold["cll_X"][:, 0], device="cuda:0")
    for s in range(36):
        U = ctrl.step(X)
        assert (ctrl.status.cpu().numpy() == 0).all(), s
        assert _relerr(U.cpu().numpy(), gold["cll_U"][:, s]) < 1e-6, s
        assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["cll_iter"][:, s]), s
        X = ctrl.plant_step(X, U)
        assert _relerr(X.cpu().numpy(), gold["cll_X"][:, s + 1]) < 1e-6, s
    assert ctrl.index == 36


def test_closed_loop_tools_batched(torch_mod):
    """tunempc/closed_loop_tools.py drivers with the Python loops as the batch axis: closed_loop_sim on the periodic
    unicycle against the oracle's rollouts, check_equivalence on the CSTR alpha sweep against single solves."""
    torch = torch_mod
    from tunempc_b200 import closed_loop_tools as clt
    ctrl, pb = _ctrl("unicycle")
    gold = load_golden("unicycle")
    log = clt.closed_loop_sim({"TUNEMPC": ctrl}, None, None, None, gold["cl_X"][:, 0], 6)
    X = torch.stack(log["x"]["TUNEMPC"], dim=1).cpu().numpy()
    U = torch.stack(log["u"]["TUNEMPC"], dim=1).cpu().numpy()
    assert _relerr(X, gold["cl_X"]) < 1e-6 and _relerr(U, gold["cl_U"]) < 1e-6
    L = torch.stack(log["l"]["TUNEMPC"], dim=1).cpu().numpy()
    z, y, u = gold["cl_X"][:, :6, 0], gold["cl_X"][:, :6, 1], gold["cl_U"][:, :, 0]
    assert _relerr(L, u ** 2 + z ** 2 + 5 * y ** 2) < 1e-6                      # examples/unicycle/main.py:82
    v = clt.reduce_rollout_stats(clt.rollout_stats(log, "TUNEMPC"))
    assert v[0] == 8 and v[1] == 6 and v[4] == 0 and abs(v[2].item() - L.sum()) < 1e-9 * L.sum()
    # alpha sweep (examples/cstr/main.py:124-134): every alpha is one instance of one batched step
    c2, pb2 = _ctrl("cstr")
    xs = pb2.wref[0, :4]
    dx = np.array([1.0 - xs[0], 0.0, 0.0, 0.0])
    alpha = np.linspace(-0.1, 1.0, 12)
    lg = clt.check_equivalence({"TUNEMPC": c2}, None, None, xs, dx, alpha)
    assert (lg["status"]["TUNEMPC"].cpu().numpy() == 0).all()
    c3, _ = _ctrl("cstr")
    for b in (0, 5, 11):
        c3.reset()
        u1 = c3.step(xs + alpha[b] * dx)
        assert _relerr(lg["u"]["TUNEMPC"][b, 0].cpu().numpy(), u1) < 1e-12
    hv = lg["h"]["TUNEMPC"].cpu().numpy()
    assert np.allclose(hv, lg["u"]["TUNEMPC"][:, :, 0].cpu().numpy() - 5.0)      # first row of h: Vdot - 5 >= 0
    cB, Vd, QK = lg["x"]["TUNEMPC"][:, :, 1].cpu().numpy(), lg["u"]["TUNEMPC"][:, :, 0].cpu().numpy(), lg["u"]["TUNEMPC"][:, :, 1].cpu().numpy()
    lref = 100 * (-cB / 5.10 + 0.1 * (1e-4 * (Vd - 14.19) ** 2 + 1e-4 * (QK + 1113.5) ** 2))   # cstr_model.py:117-129
    assert _relerr(lg["l"]["TUNEMPC"].cpu().numpy(), lref) < 1e-12


@pytest.mark.parametrize("tag,tol", [("t6", 1e-6), ("t9", 1e-9)])
def test_evaporation_golden(torch_mod, tag, tol):
    """config #3 (examples/evaporation_process): collocation integrator on the device, pure state constraints relaxed at
    stage 0, 29-row active set with non-zero reference multipliers; open loop + 5 closed-loop steps."""
    torch = torch_mod
    ctrl, pb = _ctrl("evaporation", tol=tol)
    gold = load_golden("evaporation")
    U = ctrl.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert (ctrl.status.cpu().numpy() == 0).all()
    assert _relerr(U, gold["u0_" + tag]) < 1e-6
    assert _relerr(ctrl.w_sol.cpu().numpy(), gold["w_" + tag]) < 1e-6
    lam = ctrl.lam_g.cpu().numpy()
    for b in range(lam.shape[0]):
        assert set(np.nonzero(lam[b])[0]) == set(np.nonzero(gold["lam_" + tag][b])[0]), b
    assert np.array_equal(ctrl.log["nAS"][-1].cpu().numpy(), gold["nAS_" + tag])
    assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["iter_" + tag])
    if tag == "t6":
        ctrl.reset()
        X = torch.tensor(gold["cl_X"][:, 0], device="cuda:0")
        for s in range(5):
            Uc = ctrl.step(X)
            assert _relerr(Uc.cpu().numpy(), gold["cl_U"][:, s]) < 1e-6, s
            X = ctrl.plant_step(X, Uc)
            assert _relerr(X.cpu().numpy(), gold["cl_X"][:, s + 1]) < 1e-6, s


def test_edge_cases(torch_mod):
    """ragged / tiny / empty batches, infeasible and non-finite instances inside a batch, iteration cap, reset semantics"""
    torch = torch_mod
    from oracle import reference_port as rp
    ctrl, pb = _ctrl("cstr")
    gold = load_golden("cstr")
    xs = pb.wref[0, :4]
    # batch sizes that are not multiples of the warp / CTA sizes give the same per-instance answers
    ref = None
    for B in (1, 2, 31, 33, 47):
        ctrl.reset(B)
        U = ctrl.step(torch.tensor(gold["X0"][:B], device="cuda:0")).cpu().numpy()
        assert U.shape == (B, 2) and (ctrl.status.cpu().numpy() == 0).all()
        if ref is None:
            ref = U[0]
        assert np.array_equal(U[0], ref)                                         # independent of the neighbours: bitwise
        assert _relerr(U, gold["u0_t6"][:B]) < 1e-6
    # empty batch: no launch, empty outputs, the phase index still advances (pmpc.py:415)
    ctrl.reset(0)
    U = ctrl.step(torch.empty((0, 4), dtype=torch.float64, device="cuda:0"))
    assert U.shape == (0, 2) and ctrl.index == 1
    U = ctrl.step(np.empty((0, 4)))
    assert U.shape == (0, 2) and ctrl.index == 2
    # an infeasible and a non-finite instance do not disturb their neighbours: per-instance status instead of the
    # reference's exception (CasADi conic raises on the infeasible QP; here status 2 / 4)
    X0 = gold["X0"][:6].copy()
    X0[2] = xs
    X0[2, 0] += 2.5 * (1.0 - xs[0])                                              # negative concentration: terminal set unreachable
    X0[4, 2] = np.nan
    ctrl.reset(6)
    U = ctrl.step(torch.tensor(X0, device="cuda:0")).cpu().numpy()
    st = ctrl.status.cpu().numpy()
    assert st[2] in (2, 5) and st[4] == 4 and (st[[0, 1, 3, 5]] == 0).all()          # 5: the dual active set ran out of rows before proving infeasibility
    assert _relerr(U[[0, 1, 3, 5]], gold["u0_t6"][[0, 1, 3, 5]]) < 1e-6
    oc = rp.Pmpc(pb)
    with pytest.raises(RuntimeError):
        oc.step(X0[2])                                                           # the oracle's QP solver reports infeasibility
    # iteration cap (sqp_method.py:281): status 1 after exactly max_iter iterations
    c1, _ = _ctrl("cstr", max_iter=2)
    c1.step(torch.tensor(gold["X0"][:8], device="cuda:0"))
    it = c1.log["iter"][-1].cpu().numpy()
    st = c1.status.cpu().numpy()
    assert (it <= 2).all() and ((st == 1) == (gold["iter_t6"][:8] > 2)).sum() >= 6 and (it[st == 1] == 2).all()
    # reset(): phase index and log cleared, warm start back at the reference (pmpc.py:858-865)
    ctrl.reset(3)
    a = ctrl.step(torch.tensor(gold["X0"][:3], device="cuda:0")).clone()
    ctrl.step(torch.tensor(gold["X0"][3:6], device="cuda:0"))
    assert ctrl.index == 2 and len(ctrl.log["u0"]) == 2
    ctrl.reset()
    assert ctrl.index == 0 and len(ctrl.log["u0"]) == 0
    b = ctrl.step(torch.tensor(gold["X0"][:3], device="cuda:0"))
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        ctrl.step(torch.tensor(gold["X0"][:5], device="cuda:0"))                 # batch size change without reset()
    with pytest.raises(TypeError):
        ctrl.step(torch.tensor(gold["X0"][:3], device="cuda:0", dtype=torch.float32))


def test_gauss_newton_other_configs(torch_mod):
    """hessian_approximation='gauss_newton' (pmpc.py:327-333) on the collocation and the periodic config: same solution"""
    torch = torch_mod
    for name in ("evaporation", "unicycle"):
        gold = load_golden(name)
        c, pb = _ctrl(name, hessian_approximation="gauss_newton", tol=1e-9)
        U = c.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
        assert (c.status.cpu().numpy() == 0).all()
        assert _relerr(U, gold["u0_t9"]) < 1e-6


@pytest.mark.parametrize("name", ["cstr", "evaporation"])
def test_economic_controller(torch_mod, name):
    """economic MPC (tuner.py:180-182, pmpc.py:97-107) on the device: golden outputs of the oracle, and the property the
    whole reference is about -- the tuned tracking controller is first-order equivalent to the economic one at the
    reference (closed_loop_tools.check_equivalence; paper eq. (6)): equal feedback slopes du0/dalpha as alpha -> 0."""
    torch = torch_mod
    from tunempc_b200 import closed_loop_tools as clt
    ce, pe = _ctrl(name + "_economic")
    gold = load_golden(name + "_economic")
    U = ce.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert (ce.status.cpu().numpy() == 0).all()
    assert _relerr(U, gold["u0_t6"]) < 1e-6
    assert _relerr(ce.w_sol.cpu().numpy(), gold["w_t6"]) < 1e-5
    lam = ce.lam_g.cpu().numpy()
    for b in range(lam.shape[0]):
        assert set(np.nonzero(lam[b])[0]) == set(np.nonzero(gold["lam_t6"][b])[0]), b
    c9, _ = _ctrl(name + "_economic", tol=1e-9)
    U9 = c9.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert (c9.status.cpu().numpy() == 0).all() and _relerr(U9, gold["u0_t9"]) < 1e-6
    # first-order equivalence tuned <-> economic: sweep x0 = x_ref + alpha*dx for small alpha, compare slopes
    ct, pt = _ctrl(name, tol=1e-10)
    ce2, _ = _ctrl(name + "_economic", tol=1e-10)
    xs = pt.wref[0, :pt.nx]
    dx = np.zeros(pt.nx)
    dx[-1 if name == "evaporation" else 0] = 1.0                                 # P2 (evaporation main.py:178-180) / cA
    alpha = np.array([0.0, 1e-3, 2e-3])
    lg = clt.check_equivalence({"tuned": ct, "economic": ce2}, None, None, xs, dx, alpha)
    ut = lg["u"]["tuned"][:, 0].cpu().numpy()
    ue = lg["u"]["economic"][:, 0].cpu().numpy()
    assert np.allclose(ut[0], pt.wref[0, pt.nx:], rtol=1e-9) and np.allclose(ue[0], pt.wref[0, pt.nx:], rtol=1e-9)
    st, se = (ut[1] - ut[0]) / alpha[1], (ue[1] - ue[0]) / alpha[1]
    assert np.max(np.abs(st - se)) < 1e-3 * np.max(np.abs(se)), (st, se)          # equal slopes; what remains is O(alpha)
    d1 = np.max(np.abs((ut[1] - ut[0]) - (ue[1] - ue[0])))
    d2 = np.max(np.abs((ut[2] - ut[0]) - (ue[2] - ue[0])))
    assert 3.5 < d2 / d1 < 4.5, (d1, d2)                                         # the difference is second order in alpha


@pytest.mark.parametrize("qpmode", [None, "t"])
def test_awe9_slack_formulation(torch_mod, qpmode, monkeypatch):
    """config #5 stand-in (configs.awe9: nx = 9, nu = 3, ns = 3, nsc = 3, nh = 17, N = 20, p = 40, nx_term = 7): slack variables,
    nonlinear equality rows, L1 slack cost and the bug-compatible stage-0 relaxation on the device, against the oracle"""
    torch = torch_mod
    if qpmode is not None:
        monkeypatch.setenv("TMPC_QP_MODE", qpmode)
        monkeypatch.setenv("TMPC_QP_THREAD_MIN", "1")
    ctrl, pb = _ctrl("awe9")
    gold = load_golden("awe9")
    U = ctrl.step(torch.tensor(pb.wref[0, :pb.nx][None], device="cuda:0")).cpu().numpy()
    assert ctrl.status.cpu().numpy()[0] == 0 and np.allclose(U[0], pb.wref[0, pb.nx:pb.nx + pb.nu], atol=1e-10)
    ctrl.reset()
    U = ctrl.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert U.shape == (gold["X0"].shape[0], 3)
    assert (ctrl.status.cpu().numpy() == 0).all()
    assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["iter_t6"])
    assert _relerr(U, gold["u0_t6"]) < 1e-8 and _relerr(ctrl.w_sol.cpu().numpy(), gold["w_t6"]) < 1e-8
    lam = ctrl.lam_g.cpu().numpy()
    assert _relerr(lam, gold["lam_t6"]) < 1e-6
    for b in range(lam.shape[0]):
        for k in range(pb.N):
            assert np.array_equal(lam[b][pb.g_h(k)] != 0, gold["lam_t6"][b][pb.g_h(k)] != 0), (b, k)
    for key in ("nAS", "nACtot", "nAC"):
        assert np.array_equal(ctrl.log[key][-1].cpu().numpy(), gold[key + "_t6"]), key
    assert _relerr(ctrl.log["f"][-1].cpu().numpy(), gold["f_t6"]) < 1e-8
    g = ctrl.g_sol.cpu().numpy()
    for k in range(pb.N):                                                     # g rows: h_nl(x,u) - us = 0 at the solution
        assert np.abs(g[:, pb.g_g(k)]).max() < 1e-6
    c9, _ = _ctrl("awe9", tol=1e-9)
    U9 = c9.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert (c9.status.cpu().numpy() == 0).all() and _relerr(U9, gold["u0_t9"]) < 1e-8
    assert _relerr(c9.w_sol.cpu().numpy(), gold["w_t9"]) < 1e-8
    # closed loop over the periodic reference (plant = model)
    ctrl.reset()
    X = torch.tensor(gold["cl_X"][:, 0].copy(), device="cuda:0")
    for s in range(gold["cl_U"].shape[1]):
        U = ctrl.step(X)
        assert (ctrl.status.cpu().numpy() == 0).all()
        assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["cl_iter"][:, s]), s
        assert _relerr(U.cpu().numpy(), gold["cl_U"][:, s]) < 1e-7, s
        X = ctrl.plant_step(X, U)
    assert _relerr(X.cpu().numpy(), gold["cl_X"][:, -1]) < 1e-7


def test_awe9_large_fixture(torch_mod):
    """512 seeded x0 of the config #5 stand-in against the committed oracle fixture: u0, x_1, iteration counts, f, nAS / nACtot /
    nAC and the active set of every instance"""
    torch = torch_mod
    import os
    from tunempc_b200 import configs
    ctrl, pb = _ctrl("awe9")
    L = np.load(os.path.join(os.path.dirname(__file__), "golden", "large_awe9.npz"))
    X0 = configs.sample_x0("awe9", pb, int(L["B"]), int(L["seed"]))
    U = ctrl.step(torch.tensor(X0, device="cuda:0")).cpu().numpy()
    ok = L["status"] == 0
    st = ctrl.status.cpu().numpy()
    assert np.array_equal(st[ok], L["status"][ok]), np.bincount(st)
    assert _relerr(U[ok], L["u0"][ok]) < 1e-6
    w = ctrl.w_sol.cpu().numpy()
    assert _relerr(w[ok][:, pb.ix(1)], L["x1"][ok]) < 1e-6
    assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy()[ok], L["iter"][ok])
    lam = ctrl.lam_g.cpu().numpy()
    act = np.array([np.packbits(np.array([lam[b][pb.g_h(k)][j] != 0 for k in range(pb.N) for j in range(pb.nh)], dtype=bool)) for b in range(lam.shape[0])])
    assert np.array_equal(act[ok], L["active"][ok])
    for key in ("nAS", "nACtot", "nAC"):
        assert np.array_equal(ctrl.log[key][-1].cpu().numpy()[ok], L[key][ok]), key
    assert _relerr(ctrl.log["f"][-1].cpu().numpy()[ok], L["f"][ok]) < 1e-7


def test_economic_periodic_unicycle(torch_mod):
    """economic MPC on the periodic unicycle reference (pmpc.py:97-107,709-767, p = N = 30): closed loops of the oracle"""
    torch = torch_mod
    ctrl, pb = _ctrl("unicycle_economic")
    gold = load_golden("unicycle_economic")
    X = torch.tensor(gold["X0"], device="cuda:0")
    for s in range(gold["cl_U"].shape[1]):
        U = ctrl.step(X)
        assert (ctrl.status.cpu().numpy() == 0).all()
        assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["cl_iter"][:, s])
        assert _relerr(U.cpu().numpy(), gold["cl_U"][:, s]) < 1e-8, s
        X = ctrl.plant_step(X, U)
    assert _relerr(X.cpu().numpy(), gold["cl_X"][:, -1]) < 1e-8


@pytest.mark.parametrize("name,lin_mode", [("chain", "2"), ("chain", "1"), ("dims9", "2"), ("dims9", "1"), ("dims9", None)])
def test_generic_dimensions(torch_mod, name, lin_mode, monkeypatch):
    """synthetic models beyond the reference configs' dimensions -- chain (nz = 8) and dims9 (nz = 12: the AWE config's
    dimensions, one warp-level QP per CTA): both linearisation kernels (warp-specialised with 12 / 26 consumer warps, and
    pair-per-thread) and both first-QP routes against the oracle's golden outputs"""
    torch = torch_mod
    if lin_mode:
        monkeypatch.setenv("TMPC_LIN_MODE", lin_mode)
    monkeypatch.setenv("TMPC_QP0_MIN", "2" if lin_mode == "1" else "1024")   # lin_mode None: the library's own choice
    ctrl, pb = _ctrl(name)
    gold = load_golden(name)
    U = ctrl.step(torch.tensor(gold["X0"], device="cuda:0")).cpu().numpy()
    assert (ctrl.status.cpu().numpy() == 0).all()
    assert _relerr(U, gold["u0_t6"]) < 1e-6 and _relerr(ctrl.w_sol.cpu().numpy(), gold["w_t6"]) < 1e-6
    assert np.array_equal(ctrl.log["iter"][-1].cpu().numpy(), gold["iter_t6"])
    assert np.array_equal(ctrl.log["nAS"][-1].cpu().numpy(), gold["nAS_t6"])


@pytest.mark.parametrize("name,qpmode", [("lq", None), ("cstr", None), ("cstr", "t"), ("cstr", "w"), ("evaporation", None),
                                         ("evaporation", "t"), ("unicycle", None), ("unicycle", "t")])
def test_large_fixture(torch_mod, name, qpmode, monkeypatch):
    """4096 seeded x0 per config against the committed oracle fixture (tests/golden/make_golden_large.py):
    u0 and x_1 to 1e-6 relative, identical active sets, the log outputs f / nAS / nACtot / nAC, iteration counts on the
    instances whose QPs stayed convex -- with the default kernel dispatch and with each QP kernel forced."""
    torch = torch_mod
    import os
    from test_twin import _large_compare
    from tunempc_b200 import configs
    if qpmode is not None:
        monkeypatch.setenv("TMPC_QP_MODE", qpmode)
        monkeypatch.setenv("TMPC_QP_THREAD_MIN", "1")
    ctrl, pb = _ctrl(name)
    L = np.load(os.path.join(os.path.dirname(__file__), "golden", "large_%s.npz" % name))
    n = int(L["B"])
    X0 = configs.sample_x0(name, pb, n, int(L["seed"]))
    U = ctrl.step(torch.tensor(X0, device="cuda:0")).cpu().numpy()
    lg = ctrl.log
    o = {"status": ctrl.status.cpu().numpy(), "u0": U, "x1": ctrl.w_sol[:, pb.nz:pb.nz + pb.nx].cpu().numpy(),
         "nAS": lg["nAS"][-1].cpu().numpy(), "nACtot": lg["nACtot"][-1].cpu().numpy(), "nAC": lg["nAC"][-1].cpu().numpy(),
         "f": lg["f"][-1].cpu().numpy(), "flags": lg["flags"][-1].cpu().numpy(), "iter": lg["iter"][-1].cpu().numpy()}
    nclean = _large_compare(name, pb, L, n, o, ctrl.lam_g.cpu().numpy(), "gpu")
    if name != "cstr":
        assert nclean == n                                                       # convex throughout: the oracle's iteration counts everywhere


def test_status_not_pd(torch_mod):
    """status 3 (sqp_method.py:193-201), same case as tests/test_twin.py::test_twin_status_not_pd"""
    torch = torch_mod
    pb = load_problem("lq")
    pb.H = 0.0 * pb.H
    pb.q = 0.0 * pb.q
    from tunempc_b200.pmpc import Pmpc
    ctrl = Pmpc(pb, device=0)
    ctrl.step(torch.tensor(load_golden("lq")["X0"][:8], device="cuda:0"))
    assert (ctrl.status.cpu().numpy() == 3).all()

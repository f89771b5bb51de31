"""CPU tests of the CUDA source compiled as a 1-lane host program (tests/twin) against the oracle / golden fixtures.
Checks the algorithm the kernels implement (pair-wise sensitivity integration, Riccati + dual active-set QP, filter line
search, convergence, shift) without a GPU; the same comparisons run against the real kernels in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from conftest import load_golden, load_problem


@pytest.fixture(scope="module")
def env(built):
    from oracle import reference_port as rp
    from tunempc_b200.problem import build_tables
    from twin.twin import Twin
    return rp, build_tables, Twin


def _relerr(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))


def test_twin_stage_eval_matches_oracle(env):
    rp, build_tables, Twin = env
    for name in ("lq", "cstr"):
        pb = load_problem(name)
        tw = Twin(pb, build_tables(pb))
        rng = np.random.default_rng(1)
        z = pb.wref[0] * (1 + 0.05 * rng.standard_normal((6, pb.nz))) + 0.1 * rng.standard_normal((6, pb.nz))
        a = rp.StageLib(name).F(z[:, :pb.nx], z[:, pb.nx:], 2)
        b = tw.stage_eval(z[:, :pb.nx], z[:, pb.nx:], 2)
        for x, y in zip(a, b):
            assert np.max(np.abs(x - y)) <= 1e-12 * max(1.0, np.max(np.abs(x)))


def test_twin_lq_golden(env):
    rp, build_tables, Twin = env
    pb, gold = load_problem("lq"), load_golden("lq")
    tw = Twin(pb, build_tables(pb))
    tw.reset(64)
    o = tw.step(gold["X0"])
    assert (o["status"] == 0).all() and (o["iter"] == 1).all()
    assert _relerr(o["u0"], gold["u0_t6"]) < 1e-10
    assert _relerr(o["w"], gold["w_t6"]) < 1e-9
    assert _relerr(o["lam"], gold["lam_t6"]) < 1e-8


def test_twin_cstr_golden(env):
    rp, build_tables, Twin = env
    pb, gold = load_problem("cstr"), load_golden("cstr")
    n = 48
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    tw.reset(n)
    o = tw.step(gold["X0"][:n])
    assert (o["status"] == 0).all()
    assert _relerr(o["u0"], gold["u0_t6"][:n]) < 1e-6                 # north_star tolerance: 1e-6 relative on u0
    assert _relerr(o["w"], gold["w_t6"][:n]) < 1e-5                   # both sides stop at KKT residual 1e-6
    for b in range(n):                                                # identical active sets
        assert set(np.nonzero(o["lam"][b])[0]) == set(np.nonzero(gold["lam_t6"][b])[0])
    assert np.array_equal(o["nAS"], gold["nAS_t6"][:n])
    clean = (o["flags"] & 13) == 0                                        # no convexification: same iteration path
    assert clean.any() and np.array_equal(o["iter"][clean], gold["iter_t6"][:n][clean])


def test_twin_shift_and_closed_loop(env):
    rp, build_tables, Twin = env
    pb, gold = load_problem("cstr"), load_golden("cstr")
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    st = rp.StageLib("cstr")
    n = 2
    tw.reset(n)
    x = gold["X0"][:n].copy()
    for s in range(3):
        o = tw.step(x)
        assert _relerr(o["u0"], gold["cl_U"][:n, s]) < 1e-6
        x = st.F(x, o["u0"])
        assert _relerr(x, gold["cl_X"][:n, s + 1]) < 1e-6
    # shift: compare with the oracle's own shift of the same solution
    oc = rp.Pmpc(pb)
    ws, ls = oc._shift(o["w"][0], o["lam"][0])
    assert np.array_equal(tw.W[0], ws) and np.array_equal(tw.LAM[0], ls)


def test_twin_unicycle_periodic(env):
    """config #4: periodic reference (p = N = 30, time-varying H), projected terminal constraint (nx_term = 3 < nx),
    closed loop across the period wrap (phase index 30 -> 0) with the shifted warm start."""
    rp, build_tables, Twin = env
    pb, gold = load_problem("unicycle"), load_golden("unicycle")
    assert pb.p == 30 and pb.nx_term == 3
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    tw.reset(16)
    o = tw.step(gold["X0"])
    assert (o["status"] == 0).all()
    assert np.array_equal(o["iter"], gold["iter_t6"])                  # no inequality rows: same iteration path
    assert _relerr(o["u0"], gold["u0_t6"]) < 1e-9
    assert _relerr(o["w"], gold["w_t6"]) < 1e-9
    assert _relerr(o["lam"], gold["lam_t6"]) < 1e-8
    st = rp.StageLib("unicycle")
    tw.reset(2)
    x = gold["X0"][:2].copy()
    for s in range(36):
        o = tw.step(x)
        assert (o["status"] == 0).all()
        assert _relerr(o["u0"], gold["cll_U"][:2, s]) < 1e-8, s
        assert np.array_equal(o["iter"], gold["cll_iter"][:2, s]), s
        x = st.F(x, o["u0"])
        assert _relerr(x, gold["cll_X"][:2, s + 1]) < 1e-8, s


def test_twin_shared_first_qp(env):
    """first QP after reset(): the table route (tm_qp0_*: one tabulated parametric QP, working-set iteration per
    instance) against the per-instance Riccati + dual active-set route, and against the oracle's golden answers."""
    rp, build_tables, Twin = env
    pb, gold = load_problem("cstr"), load_golden("cstr")
    n = 48
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    tw.reset(n)
    a = tw.step(gold["X0"][:n])
    tw.reset(n)
    b = tw.step(gold["X0"][:n], shared_first_qp=True)
    assert (b["status"] == 0).all() and np.array_equal(a["iter"], b["iter"]) and np.array_equal(a["nAS"], b["nAS"])
    assert b["counters"][7] < a["counters"][7]                            # fewer Riccati sweeps
    assert _relerr(a["u0"], b["u0"]) < 1e-9 and _relerr(a["w"], b["w"]) < 1e-9
    assert _relerr(b["u0"], gold["u0_t6"][:n]) < 1e-6
    for i in range(n):
        assert set(np.nonzero(b["lam"][i])[0]) == set(np.nonzero(gold["lam_t6"][i])[0])
    # the QP itself: stop after one SQP iteration and compare step and multipliers of the two routes
    pb1 = load_problem("cstr")
    pb1.max_iter = 1
    t1 = Twin(pb1, build_tables(pb1), rho=3e7, al_gamma=1e3)
    t1.reset(n)
    a1 = t1.step(gold["X0"][:n])
    t1.reset(n)
    b1 = t1.step(gold["X0"][:n], shared_first_qp=True)
    assert np.array_equal(a1["nAS"], b1["nAS"])
    assert _relerr(a1["w"], b1["w"]) < 1e-9 and _relerr(a1["lam"], b1["lam"]) < 1e-6
    for i in range(n):
        assert set(np.nonzero(a1["lam"][i])[0]) == set(np.nonzero(b1["lam"][i])[0])     # exact zeros off the working set


def test_twin_evaporation_collocation(env):
    """config #3: implicit (Radau collocation) integrator with IFT sensitivities, pure state constraints relaxed at
    stage 0, 29 active rows at the solution (reference on the bound X2 = 25), reduced-space convexification mask."""
    rp, build_tables, Twin = env
    pb, gold = load_problem("evaporation"), load_golden("evaporation")
    assert pb.h_x_idx == [0, 1, 2] and pb.lam_h_ref[0, 0] < 0
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    rng = np.random.default_rng(3)
    z = pb.wref[0] * (1 + 0.03 * rng.standard_normal((5, pb.nz)))
    a = rp.StageLib("evaporation").F(z[:, :pb.nx], z[:, pb.nx:], 2)           # full-tensor IFT on the stage states
    b = tw.stage_eval(z[:, :pb.nx], z[:, pb.nx:], 2)                          # pair-wise IFT on the stage derivatives
    for x, y in zip(a, b):
        assert np.max(np.abs(x - y)) <= 1e-11 * max(1.0, np.max(np.abs(x)))
    n = 24
    tw.reset(n)
    o = tw.step(gold["X0"])
    assert (o["status"] == 0).all() and (o["flags"] & 1 == 0).all()
    assert np.array_equal(o["iter"], gold["iter_t6"]) and np.array_equal(o["nAS"], gold["nAS_t6"])
    assert _relerr(o["u0"], gold["u0_t6"]) < 1e-9 and _relerr(o["w"], gold["w_t6"]) < 1e-9
    for i in range(n):
        assert set(np.nonzero(o["lam"][i])[0]) == set(np.nonzero(gold["lam_t6"][i])[0])
    st = rp.StageLib("evaporation")
    tw.reset(4)
    x = gold["cl_X"][:, 0].copy()
    for s in range(5):
        o = tw.step(x)
        assert _relerr(o["u0"], gold["cl_U"][:, s]) < 1e-8, s
        x = st.F(x, o["u0"])


def test_twin_economic_controller(env):
    """create_mpc('economic') (tuner.py:180-182, pmpc.py:97-107): stage cost l(x,u) of the model card, exact Hessian,
    non-zero dynamics multipliers in the dual reference.  Same converged points and active sets as the oracle."""
    rp, build_tables, Twin = env
    for name in ("cstr", "evaporation"):
        pb, gold = load_problem(name + "_economic"), load_golden(name + "_economic")
        assert pb.mpc_type == "economic" and np.abs(pb.lam_dyn_ref).max() > 0
        n = gold["X0"].shape[0]
        tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
        tw.reset(n)
        o = tw.step(gold["X0"])
        assert (o["status"] == 0).all()
        assert _relerr(o["u0"], gold["u0_t6"]) < 1e-6 and _relerr(o["w"], gold["w_t6"]) < 1e-5
        for b in range(n):
            assert set(np.nonzero(o["lam"][b])[0]) == set(np.nonzero(gold["lam_t6"][b])[0]), (name, b)
        clean = (o["flags"] & 13) == 0
        assert np.array_equal(o["iter"][clean], gold["iter_t6"][clean])
        # at the reference the economic controller returns the reference input in one iteration (P1)
        tw.reset(1)
        o = tw.step(pb.wref[0, :pb.nx][None])
        assert o["iter"][0] == 1 and np.allclose(o["u0"][0], pb.wref[0, pb.nx:], rtol=1e-9)


def test_twin_economic_periodic_unicycle(env):
    """economic MPC on a periodic reference (pmpc.py:97-107,709-767 with p = N = 30): closed loops of the oracle, phase-indexed
    dual reference with non-zero dynamics multipliers and the projected terminal multiplier"""
    rp, build_tables, Twin = env
    pb, gold = load_problem("unicycle_economic"), load_golden("unicycle_economic")
    assert pb.mpc_type == "economic" and pb.p == 30 and np.abs(pb.lam_dyn_ref).max() > 0
    st = rp.StageLib("unicycle")
    tw = Twin(pb, build_tables(pb))
    tw.reset(gold["X0"].shape[0])
    x = gold["X0"].copy()
    for s in range(gold["cl_U"].shape[1]):
        o = tw.step(x)
        assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["cl_iter"][:, s])
        assert _relerr(o["u0"], gold["cl_U"][:, s]) < 1e-9, s
        x = st.F(x, o["u0"])
    assert _relerr(x, gold["cl_X"][:, -1]) < 1e-9


def test_twin_awe9_slack_formulation(env):
    """config #5 stand-in (configs.awe9): slack variables us / usc, nonlinear equality rows g, L1 slack cost, bug-compatible
    stage-0 relaxation, p = 40 periodic reference, 7-row projected terminal constraint -- the device routines against the
    oracle's dense restatement (pmpc.py:217-294,338-339,709-721)"""
    rp, build_tables, Twin = env
    pb, gold = load_problem("awe9"), load_golden("awe9")
    tw = Twin(pb, build_tables(pb))
    tw.reset(1)
    o = tw.step(pb.wref[0, :pb.nx][None])                                   # P1: step(x_ref) = u_ref in one iteration
    assert o["status"][0] == 0 and o["iter"][0] == 1 and np.allclose(o["u0"][0], pb.wref[0, pb.nx:pb.nx + pb.nu], atol=1e-10)
    n = 8
    tw.reset(n)
    o = tw.step(gold["X0"][:n])
    assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["iter_t6"][:n])
    assert _relerr(o["u0"], gold["u0_t6"][:n]) < 1e-9 and _relerr(o["w"], gold["w_t6"][:n]) < 1e-9
    assert _relerr(o["lam"], gold["lam_t6"][:n]) < 1e-7
    for key in ("nAS", "nACtot", "nAC"):
        assert np.array_equal(o[key], gold[key + "_t6"][:n]), key
    assert _relerr(o["f"], gold["f_t6"][:n]) < 1e-9
    for b in range(n):                                                      # identical active sets (inequality rows)
        for k in range(pb.N):
            assert np.array_equal(o["lam"][b][pb.g_h(k)] != 0, gold["lam_t6"][b][pb.g_h(k)] != 0), (b, k)
    # closed loop over the periodic reference: phase tables, warm-start shift with slacks (pmpc.py:867-906)
    st = rp.StageLib("awe9")
    nb = 3
    tw.reset(nb)
    x = gold["cl_X"][:nb, 0].copy()
    for s in range(gold["cl_U"].shape[1]):
        o = tw.step(x)
        assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["cl_iter"][:nb, s]), s
        assert _relerr(o["u0"], gold["cl_U"][:nb, s]) < 1e-8, s
        x = st.F(x, o["u0"])


def test_twin_generic_dimensions_chain(env):
    """synthetic nx = 6, nu = 2 model (configs.chain, not in the reference): nothing in the device code is tied to the
    dimensions of the four reference configs"""
    rp, build_tables, Twin = env
    pb, gold = load_problem("chain"), load_golden("chain")
    assert pb.nz == 8
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    n = gold["X0"].shape[0]
    tw.reset(n)
    o = tw.step(gold["X0"])
    assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["iter_t6"]) and np.array_equal(o["nAS"], gold["nAS_t6"])
    assert _relerr(o["u0"], gold["u0_t6"]) < 1e-9 and _relerr(o["w"], gold["w_t6"]) < 1e-9
    tw.reset(n)
    o2 = tw.step(gold["X0"], shared_first_qp=True)
    assert _relerr(o2["u0"], gold["u0_t6"]) < 1e-9 and np.array_equal(o2["iter"], gold["iter_t6"])


def test_twin_awe_dimensions_dims9(env):
    """synthetic stand-in with the dimensions of the AWE config (nx = 9, nu = 3, 14 constraint rows, N = 20, 7-row projected
    terminal constraint; configs.dims9): state-only rows relaxed at stage 0, mixed rows, 280 inequality rows"""
    rp, build_tables, Twin = env
    pb, gold = load_problem("dims9"), load_golden("dims9")
    assert (pb.nx, pb.nu, pb.nh, pb.nx_term, pb.n_w, pb.n_g) == (9, 3, 14, 7, 249, 476)
    tw = Twin(pb, build_tables(pb), rho=3e7, al_gamma=1e3)
    n = gold["X0"].shape[0]
    tw.reset(n)
    o = tw.step(gold["X0"], shared_first_qp=True)
    assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["iter_t6"]) and np.array_equal(o["nAS"], gold["nAS_t6"])
    assert _relerr(o["u0"], gold["u0_t6"]) < 1e-9 and _relerr(o["w"], gold["w_t6"]) < 1e-9
    ineq = np.concatenate([np.arange(pb.g_h(k).start, pb.g_h(k).stop) for k in range(pb.N)])
    for b in range(n):                    # active sets = inequality rows with a non-zero multiplier (sqp_method.py:417-423)
        assert set(np.nonzero(o["lam"][b][ineq])[0]) == set(np.nonzero(gold["lam_t6"][b][ineq])[0])
    assert _relerr(o["lam"], gold["lam_t6"]) < 1e-8


def _large_compare(name, pb, L, n, o, lam, tag):
    """shared by the CPU twin test (subset) and the GPU test (all 4096): results vs the committed oracle fixture"""
    nI = pb.N * pb.nh
    st = np.asarray(L["status"][:n])
    ok = st == 0
    assert ok.all(), "oracle fixture holds non-converged instances"
    assert (o["status"] == 0).all(), (tag, np.bincount(o["status"]))
    assert _relerr(o["u0"], L["u0"][:n]) < 1e-6                        # north_star: 1e-6 relative on u0
    assert _relerr(o["x1"], L["x1"][:n]) < 1e-6                        # ... and on the predicted trajectory
    if nI:
        act = np.stack([lam[:, pb.g_h(k)] != 0 for k in range(pb.N)], axis=1).reshape(n, nI)
        actg = np.unpackbits(L["active"][:n], axis=1)[:, :nI].astype(bool)
        assert np.array_equal(act, actg), "active sets differ"          # identical active sets
    assert np.array_equal(o["nAS"], L["nAS"][:n])
    assert np.array_equal(o["nACtot"], L["nACtot"][:n])               # active-set changes w.r.t. the initial guess (sqp_method.py:203-205)
    assert np.array_equal(o["nAC"], L["nAC"][:n])                     # stage-0 changes w.r.t. the reference multipliers (pmpc.py:840-856)
    assert _relerr(o["f"], L["f"][:n]) < 1e-5                         # objective at the returned point (both stop at KKT residual 1e-6)
    clean = (o["flags"] & 13) == 0                                     # convex QPs throughout: the oracle's iteration path
    assert clean.any() and np.array_equal(o["iter"][clean], L["iter"][:n][clean])
    return int(clean.sum())


@pytest.mark.parametrize("name,n", [("lq", 4096), ("evaporation", 384), ("unicycle", 192), ("cstr", 256)])
def test_twin_large_fixture(env, name, n):
    """the committed 4096-instance oracle fixtures (tests/golden/make_golden_large.py); the CPU suite checks a
    prefix with the sequential twin, the GPU suite all of them (test_gpu_parity.py::test_large_fixture)."""
    rp, build_tables, Twin = env
    from tunempc_b200 import configs
    pb = load_problem(name)
    L = np.load(os.path.join(os.path.dirname(__file__), "golden", "large_%s.npz" % name))
    X0 = configs.sample_x0(name, pb, int(L["B"]), int(L["seed"]))[:n]
    tw = Twin(pb, build_tables(pb))
    tw.reset(n)
    o = tw.step(X0)
    o["x1"] = o["w"][:, pb.nz:pb.nz + pb.nx]
    _large_compare(name, pb, L, n, o, o["lam"], "twin")


def test_twin_status_not_pd(env):
    """status 3 (sqp_method.py:193-201): a zero tracking Hessian leaves the reduced Hessian singular -- the oracle's
    post-solve check fails, the device reports TMPC_NOT_PD"""
    rp, build_tables, Twin = env
    pb, gold = load_problem("lq"), load_golden("lq")
    pb.H = 0.0 * pb.H
    pb.q = 0.0 * pb.q
    oc = rp.Pmpc(pb)
    oc.reset()
    oc.step(gold["X0"][0])
    assert oc.log["status"][-1] == 3
    tw = Twin(pb, build_tables(pb))
    tw.reset(4)
    o = tw.step(gold["X0"][:4])
    assert (o["status"] == 3).all()
    # and the check passes where the oracle's passes: the tuned problems converge with status 0 (every other test)

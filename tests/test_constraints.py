"""tunempc_b200/constraints.py -- the product-side path-constraint reformulations (tunempc/preprocessing.py:35-155 in matrix form):
the reference's own unit-test cases and known answers (test/test_processing.py:74-104, 147-184), the softened evaporation
controller against the oracle (CPU twin), and the Pmpc constructor path a `slack_flag` option takes."""
import numpy as np
import sympy as sp

from conftest import load_golden, load_problem
from tunempc_b200 import constraints


def test_split_mixed_constraints_known_answers():                       # test_processing.py:74-104
    x = sp.symbols("x0:1")
    u = sp.symbols("u0:2")
    C, c, gnl, where = constraints.split_path_constraints(x, u, [x[0] + u[0], u[1], x[0] ** 2 * u[0]])
    assert where == [("h", 0), ("h", 1), ("g", 0)] and len(gnl) == 1 and C.shape == (3, 4)
    z = np.array([1.0, 0.0, 2.0, 3.0])                                  # x = 1, u = (0, 2), us = 3
    assert (C @ z + c).tolist() == [1.0, 2.0, 3.0]                      # :103  h = [x+u0, u1, us]
    g = float(gnl[0].subs({x[0]: 1.0, u[0]: 0.0, u[1]: 2.0})) - z[3]
    assert g == -3.0                                                    # :104  g = x^2 u0 - us
    assert constraints.state_only_rows(x, u, gnl) == [] and constraints.state_only_rows(x, u, [x[0] ** 2]) == [0]
    C2, c2, gnl2, _ = constraints.split_path_constraints(x, u, [x[0] + u[0], u[1]])      # :51-72: nothing to slack
    assert gnl2 == [] and C2.shape == (2, 3)


def test_soften_rows_known_answers():                                   # test_processing.py:106-184
    C = np.array([[1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])                    # h = [x + u0, u1]
    c = np.zeros(2)
    Cs, cs, scost, rows = constraints.soften_rows(C, c, np.zeros((2, 2)), "active")      # :125-145 no active row: unchanged
    assert rows == [] and Cs.shape == (2, 3) and scost.size == 0
    lam = np.zeros((2, 2))
    lam[0, 0] = -5.0
    Cs, cs, scost, rows = constraints.soften_rows(C, c, lam, "active")                   # :147-184
    assert rows == [0] and Cs.shape == (3, 4)
    assert (Cs @ np.array([1.0, 0.0, 2.0, 3.0]) + cs).tolist() == [4.0, 2.0, 3.0]        # :183  h = [x+u0+usc, u1, usc]
    assert scost.tolist() == [5000.0]                                                    # :184  1e3 * max(-lam)
    Ca, _, sa, ra = constraints.soften_rows(C, c, lam, "all")
    assert ra == [0, 1] and Ca.shape == (4, 5) and sa.tolist() == [5000.0, 0.0]
    assert constraints.soften_rows(C, c, lam, "none")[3] == []


def test_softened_problem_from_constructor_arguments():
    """what Tuner.create_mpc(..., opts={'slack_flag': 'active'}) hands to Pmpc (tuner.py:171-177): the committed fixture"""
    from tunempc_b200 import configs
    from tunempc_b200.pmpc import problem_from_reference_args
    pb, ps = load_problem("evaporation"), load_problem("evaporation_sc1")
    Cs, cs, scost, rows = constraints.soften_rows(pb.C, pb.c, pb.lam_h_ref, "active")
    assert rows == [0] and np.isclose(scost[0], 1e3 * 60.29091421, rtol=1e-6)
    model = configs.CONFIGS["evaporation_sc1"]()["model"]
    assert model.name == "evaporation_sc1" and model.nsc == 1 and constraints.soft_model(model, 1) is model
    card = {"f": model, "h": (Cs, cs), "scost": scost, "vars": {"x": model.x, "u": model.u, "usc": [0]}}
    p2 = problem_from_reference_args(pb.N, card, "tracking", {"x": [pb.wref[0, :2]], "u": [pb.wref[0, 2:]]},
                                     {"H": list(pb.H), "q": list(pb.q)}, {"dyn": [np.zeros(2)], "h": list(pb.lam_h_ref)},
                                     {"A": pb.S_A, "B": pb.S_B}, {"p_operator": pb.term_idx})
    assert (p2.nz, p2.nh, p2.nsc, p2.n_w, p2.n_g) == (5, 6, 1, 152, 244) and p2.relax0 == [1, 2]
    for fld in ("C", "c", "wref", "H", "q", "scost", "lam_h_ref"):
        assert np.array_equal(getattr(p2, fld), getattr(ps, fld)), fld


def test_twin_softened_evaporation(built):
    """soft constraints on the device routines (CPU twin) against the oracle: initial states below the softened bound X2 >= 25 are
    absorbed by the slack (usc_0 = 0.3, 0.8), same iteration counts, closed loops"""
    from oracle import reference_port as rp
    from tunempc_b200.problem import build_tables
    from twin.twin import Twin
    ps, gold = load_problem("evaporation_sc1"), load_golden("evaporation_sc1")
    tw = Twin(ps, build_tables(ps))
    n = gold["X0"].shape[0]
    tw.reset(n)
    o = tw.step(gold["X0"])
    rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))
    assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["iter_t6"]) and np.array_equal(o["nAS"], gold["nAS_t6"])
    assert rel(o["u0"], gold["u0_t6"]) < 1e-9 and rel(o["w"], gold["w_t6"]) < 1e-9
    assert np.isclose(o["w"][2][ps.iusc(0)][0], 0.3, atol=1e-9) and np.isclose(o["w"][5][ps.iusc(0)][0], 0.8, atol=1e-9)
    for b in range(n):
        for k in range(1, ps.N):            # stage 0: with x_0 on the bound the split between the row and usc >= 0 is not unique
            assert np.array_equal(o["lam"][b][ps.g_h(k)] != 0, gold["lam_t6"][b][ps.g_h(k)] != 0), (b, k)
    st = rp.StageLib("evaporation")
    tw.reset(2)
    x = gold["cl_X"][:, 0].copy()
    for s in range(gold["cl_U"].shape[1]):
        o = tw.step(x)
        assert (o["status"] == 0).all() and np.array_equal(o["iter"], gold["cl_iter"][:, s]), s
        assert rel(o["u0"], gold["cl_U"][:, s]) < 1e-8, s
        x = st.F(x, o["u0"])

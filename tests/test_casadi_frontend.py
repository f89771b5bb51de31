"""CasADi front-end (tunempc_b200/casadi_frontend.py): skipped where CasADi is not importable (the build container)."""
import numpy as np
import pytest

ca = pytest.importorskip("casadi")


def test_cstr_from_casadi_matches_the_sympy_card():
    """examples/cstr/cstr_model.py written with CasADi SX -> OdeModel through the bridge: same right-hand side and
    Jacobian as the sympy model card, same linear constraint rows"""
    import sympy as sp
    from tunempc_b200 import casadi_frontend as cf, configs, modelgen
    card = configs.cstr()
    m_ref = card["model"]
    x = ca.SX.sym("x", 4)
    u = ca.SX.sym("u", 2)
    k10, k20, k30, E1, E2, E3 = 1.287e12, 1.287e12, 9.043e9, -9758.3, -9758.3, -8560.0
    DH_AB, DH_BC, DH_AD, rho, Cp, kw, AR, VR, mK, CPK, cA0, th0 = 4.2, -11.0, -41.85, 0.9342, 3.01, 4032.0, 0.215, 10.0, 5.0, 2.0, 5.10, 104.9
    k1 = k10 * ca.exp(E1 / (x[2] + 273.15)); k2 = k20 * ca.exp(E2 / (x[2] + 273.15)); k3 = k30 * ca.exp(E3 / (x[2] + 273.15))
    xdot = ca.vertcat((u[0] * (cA0 - x[0]) - k1 * x[0] - k3 * x[0] * x[0]) / 3600,
                      (-u[0] * x[1] + k1 * x[0] - k2 * x[1]) / 3600,
                      (u[0] * (th0 - x[2]) - 1.0 / (rho * Cp) * (k1 * x[0] * DH_AB + k2 * x[1] * DH_BC + k3 * x[0] * x[0] * DH_AD)
                       + kw * AR / (rho * Cp * VR) * (x[3] - x[2])) / 3600,
                      (1.0 / (mK * CPK) * (u[1] + kw * AR * (x[2] - x[3]))) / 3600)
    f = ca.Function("f", [x, u], [xdot])
    h = ca.Function("h", [x, u], [ca.vertcat(u[0] - 5.0, 35.0 - u[0], u[1] + 9000.0, -u[1])])
    m = cf.model_from_casadi("cstr_ca", f, rk_steps=20, tf=20.0)
    fa, ja = modelgen.lambdify_ode(m)
    fb, jb = modelgen.lambdify_ode(m_ref)
    z = card["w_guess"] * (1 + 0.01 * np.arange(6))
    assert np.allclose(np.array(fa(z), dtype=float), np.array(fb(z), dtype=float), rtol=1e-12)
    assert np.allclose(np.array(ja(z), dtype=float), np.array(jb(z), dtype=float), rtol=1e-12)
    C, c = cf.linear_constraints_from_casadi(h, 4, 2)
    assert np.array_equal(C, card["C"]) and np.array_equal(c, card["c"])
    hn = ca.Function("hn", [x, u], [x[0] * u[0]])
    with pytest.raises(NotImplementedError):
        cf.linear_constraints_from_casadi(hn, 4, 2)
    # nonlinear rows go through the slack reformulation (preprocessing.py:78-118): the reference's own test case (test_processing.py:74-104)
    hm = ca.Function("hm", [x, u], [ca.vertcat(x[0] + u[0], u[1], x[0] ** 2 * u[0])])
    m2, C2, c2, where = cf.constraints_from_casadi(hm, m)
    assert where == [("h", 0), ("h", 1), ("g", 0)] and m2.ns == 1 and C2.shape == (3, 7)
    zz = np.zeros(7); zz[0] = 1.0; zz[5] = 2.0; zz[6] = 3.0                      # x0 = 1, u = (0, 2), us = 3
    assert (C2 @ zz + c2).tolist() == [1.0, 2.0, 3.0]

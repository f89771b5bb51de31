"""Native offline tuning (host, run once): supplies (wref, lam_ref, S, Hc, q) to the batched controller.

The reference computes these with IPOPT + an active-set SQP (`tunempc/pocp.py:205-362`) and a picos SDP
(`tunempc/convexifier.py:36-163`), none of which exist in this image.  This module is the minimal replacement
needed to *feed* the hot path; it is NOT on the hot path:

  solve_steady_state   p = 1 OCP  min l(x,u) s.t. F(x,u) = x, C z + c >= 0      (pocp.py:78-203 with N = 1)
  sensitivities        S = {A, B, C, H, q}                                         (pocp.py:261-362)
  convexify_dare       closed-form "type A" tuning: dH(dP) with dP = P_dare - eps X  (convexifier.py:165-211 is the
                       dH(dP) map; the SDP that picks dP is replaced by a DARE + Lyapunov construction)

`stage_F(x,u,order)` is any callable returning (xf, S[, T]) for one stage -- the C-ABI library's host evaluation of
the generated model in the product, the oracle's in tests.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt
import sympy as sp


def lambdify_cost(model, cost):
    z = list(model.x) + list(model.u)
    g = [sp.diff(cost, s) for s in z]
    Hs = [[sp.diff(gi, s) for s in z] for gi in g]
    l_f = sp.lambdify([z], cost, "numpy")
    g_f = sp.lambdify([z], g, "numpy")
    H_f = sp.lambdify([z], Hs, "numpy")
    return (lambda v: float(l_f(v)), lambda v: np.array(g_f(v), dtype=np.float64),
            lambda v: np.array(H_f(v), dtype=np.float64))


def _single(stage_F):
    """adapt a batched stage evaluator ((n,nx),(n,nu),order) to one stage, squeezing the batch axis."""
    def F1(x, u, order):
        out = stage_F(np.asarray(x, dtype=np.float64)[None, :], np.asarray(u, dtype=np.float64)[None, :], order)
        return out[0] if order == 0 else tuple(o[0] for o in out)
    return F1


def solve_steady_state(stage_F, cost_funs, C, c, z0, nx, tol=1e-12, lam_tresh=1e-8):
    """returns z*, lam_dyn (nx), lam_h (nh) in CasADi sign convention  grad l + J' lam = 0, g = [F(x,u) - x ; h]."""
    stage_F = _single(stage_F)
    l_f, g_f, H_f = cost_funs
    nz = len(z0)
    nh = C.shape[0]

    def dyn(z):
        xf, S = stage_F(z[:nx], z[nx:], 1)
        return xf - z[:nx], S - np.hstack([np.eye(nx), np.zeros((nx, nz - nx))])

    cons = [{"type": "eq", "fun": lambda z: dyn(z)[0], "jac": lambda z: dyn(z)[1]}]
    if nh:
        cons.append({"type": "ineq", "fun": lambda z: C @ z + c, "jac": lambda z: C})
    scale = np.maximum(np.abs(z0), 1.0)
    res = sopt.minimize(lambda y: l_f(y * scale), z0 / scale, jac=lambda y: g_f(y * scale) * scale,
                        constraints=[{"type": cn["type"], "fun": (lambda y, cn=cn: cn["fun"](y * scale)),
                                      "jac": (lambda y, cn=cn: cn["jac"](y * scale) * scale[None, :])} for cn in cons],
                        method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
    z = res.x * scale
    # active-set Newton polish on the KKT system (what the reference's SQP re-solve does, pocp.py:255-257)
    act = [i for i in range(nh) if (C @ z + c)[i] < 1e-6 * max(1.0, abs(c[i]))]
    lam_d = np.zeros(nx)
    lam_a = np.zeros(len(act))
    if True:   # multiplier estimate from stationarity (a linear economic cost has no Hessian of its own: with lam = 0 the
        # first KKT matrix would be singular)
        _, Jd0 = dyn(z)
        J0 = np.vstack([Jd0, C[act] if act else np.zeros((0, nz))])
        le = np.linalg.lstsq(J0.T, -g_f(z), rcond=None)[0]
        lam_d, lam_a = le[:nx].copy(), le[nx:].copy()
    for it in range(50):
        r_d, Jd = dyn(z)
        Ja = C[act] if act else np.zeros((0, nz))
        Jall = np.vstack([Jd, Ja])
        xf, S, T = stage_F(z[:nx], z[nx:], 2)
        Hl = H_f(z) + np.einsum("a,aij->ij", lam_d, T)
        grad = g_f(z) + Jall.T @ np.concatenate([lam_d, lam_a])
        res_v = np.concatenate([grad, r_d, (C @ z + c)[act]])
        if np.linalg.norm(res_v / np.concatenate([np.ones(nz), np.ones(nx), np.ones(len(act))]), np.inf) < tol and it > 0:
            break
        m = Jall.shape[0]
        K = np.block([[Hl, Jall.T], [Jall, np.zeros((m, m))]])
        step = np.linalg.solve(K, -res_v)
        z = z + step[:nz]
        lam_d = lam_d + step[nz:nz + nx]
        lam_a = lam_a + step[nz + nx:]
    lam_h = np.zeros(nh)
    for j, i in enumerate(act):
        lam_h[i] = lam_a[j]
    lam_h[np.abs(lam_h) < lam_tresh] = 0.0
    if np.any(lam_h > 0):
        raise RuntimeError("steady state: wrong-signed multiplier, active set guess failed")
    return z, lam_d, lam_h


def lambdify_gnl(model):
    """numpy callables (value (ns,), Jacobian (ns, nx+nu), second derivatives (ns, nx+nu, nx+nu)) of the model card's nonlinear
    path constraints h_nl(x,u) >= 0"""
    z = list(model.x) + list(model.u)
    g = sp.Matrix([sp.sympify(e) for e in model.gnl])
    ns = len(model.gnl)
    fv = sp.lambdify([z], g, "numpy")
    fj = sp.lambdify([z], g.jacobian(z), "numpy")
    fh = [sp.lambdify([z], sp.hessian(e, z), "numpy") for e in g]
    return (lambda v: np.asarray(fv(list(v)), dtype=np.float64).ravel(),
            lambda v: np.asarray(fj(list(v)), dtype=np.float64).reshape(ns, len(z)),
            lambda v: np.array([np.asarray(h(list(v)), dtype=np.float64) for h in fh]).reshape(ns, len(z), len(z)))


def solve_steady_state_slack(stage_F, cost_funs, gnl_funs, C, c, z0, nx, ns, tol=1e-12, lam_tresh=1e-8):
    """Steady-state OCP in the slack form of tunempc/preprocessing.py:78-118 (what `Tuner(f, l, h, 1).solve_ocp` solves when h has
    nonlinear rows):  min l(x,u)  s.t.  F(x,u) = x,  g = h_nl(x,u) - us = 0,  C (x,u,us) + c >= 0  (the last ns rows of C are us >= 0).
    Solved in (x,u) with h_nl as inequalities, then polished by an active-set Newton; us = h_nl(x,u) afterwards.
    Returns z (nx+nu+ns), lam_dyn (nx), lam_g (ns), lam_h (nh) in CasADi sign; lam_g equals the multiplier of the row us >= 0
    (stationarity in us: -lam_g + lam_(us>=0) = 0)."""
    stage_F = _single(stage_F)
    l_f, g_f, H_f = cost_funs
    gv, gj, gh = gnl_funs
    nzm = len(z0)
    nh = C.shape[0]
    n_aff = nh - ns
    Ca, ca_ = C[:n_aff, :nzm], c[:n_aff]

    def dyn(z):
        xf, S = stage_F(z[:nx], z[nx:], 1)
        return xf - z[:nx], S - np.hstack([np.eye(nx), np.zeros((nx, nzm - nx))])

    def ineq(z):
        return np.concatenate([Ca @ z + ca_, gv(z)])

    def ineq_jac(z):
        return np.vstack([Ca, gj(z)])

    scale = np.maximum(np.abs(z0), 1.0)
    res = sopt.minimize(lambda y: l_f(y * scale), z0 / scale, jac=lambda y: g_f(y * scale) * scale,
                        constraints=[{"type": "eq", "fun": lambda y: dyn(y * scale)[0], "jac": lambda y: dyn(y * scale)[1] * scale[None, :]},
                                     {"type": "ineq", "fun": lambda y: ineq(y * scale), "jac": lambda y: ineq_jac(y * scale) * scale[None, :]}],
                        method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
    z = res.x * scale
    hv = ineq(z)
    act = [i for i in range(nh) if hv[i] < 1e-6]
    _, Jd0 = dyn(z)
    Jall = np.vstack([Jd0, ineq_jac(z)[act] if act else np.zeros((0, nzm))])
    le = np.linalg.lstsq(Jall.T, -g_f(z), rcond=None)[0]
    lam_d, lam_a = le[:nx].copy(), le[nx:].copy()
    for it in range(50):
        r_d, Jd = dyn(z)
        Jin = ineq_jac(z)
        Ja = Jin[act] if act else np.zeros((0, nzm))
        Jall = np.vstack([Jd, Ja])
        xf, S, T = stage_F(z[:nx], z[nx:], 2)
        Hl = H_f(z) + np.einsum("a,aij->ij", lam_d, T)
        G2 = gh(z)
        for j, i in enumerate(act):
            if i >= n_aff:
                Hl = Hl + lam_a[j] * G2[i - n_aff]
        grad = g_f(z) + Jall.T @ np.concatenate([lam_d, lam_a])
        res_v = np.concatenate([grad, r_d, ineq(z)[act]])
        if np.linalg.norm(res_v, np.inf) < tol and it > 0:
            break
        m = Jall.shape[0]
        K = np.block([[Hl, Jall.T], [Jall, np.zeros((m, m))]])
        step = np.linalg.solve(K, -res_v)
        z = z + step[:nzm]
        lam_d = lam_d + step[nzm:nzm + nx]
        lam_a = lam_a + step[nzm + nx:]
    lam_in = np.zeros(nh)
    for j, i in enumerate(act):
        lam_in[i] = lam_a[j]
    lam_in[np.abs(lam_in) < lam_tresh] = 0.0
    if np.any(lam_in > 0):
        raise RuntimeError("steady state: wrong-signed multiplier, active set guess failed")
    us = gv(z)
    us[np.abs(us) < 1e-12] = 0.0
    lam_g = lam_in[n_aff:].copy()                       # multiplier of g = that of us >= 0
    return np.concatenate([z, us]), lam_d, lam_g, lam_in


def sensitivities_slack(stage_F, cost_funs, gnl_funs, C, w, lam_d, lam_g, lam_h, nx, nu, ns):
    """S of pocp.py:261-362 for p = 1 in the slack form: stage variables (x, u, us); B carries zero columns for us, H has no
    curvature in us, the rows g = h_nl - us (always active) join the active rows of h in C_As."""
    _, _, H_f = cost_funs
    gv, gj, gh = gnl_funs
    stage_F = _single(stage_F)
    nzm, nzr = nx + nu, nx + nu + ns
    z = w[:nzm]
    xf, S1, T = stage_F(z[:nx], z[nx:], 2)
    H = np.zeros((nzr, nzr))
    Hxu = H_f(z) + np.einsum("a,aij->ij", lam_d, T) + np.einsum("a,aij->ij", lam_g, gh(z))
    H[:nzm, :nzm] = 0.5 * (Hxu + Hxu.T)
    B = np.hstack([S1[:, nx:], np.zeros((nx, ns))])
    Jg = np.hstack([gj(z), -np.eye(ns)])
    q = -(lam_h @ C)                                       # pocp.py:357-360 (the tuned controller's dual reference has lam_g = 0)
    act = [i for i in range(C.shape[0]) if lam_h[i] != 0]
    C_As = np.vstack([Jg] + ([C[act]] if act else []))
    return {"A": [S1[:, :nx].copy()], "B": [B], "C": [C.copy()], "H": [H], "q": [q], "C_As": [C_As], "G": [Jg]}


def sensitivities(stage_F, cost_funs, C, z, lam_d, lam_h, nx):
    """S of pocp.py:261-362 for p = 1 (lists of length 1, as the reference returns)."""
    _, _, H_f = cost_funs
    stage_F = _single(stage_F)
    xf, S1, T = stage_F(z[:nx], z[nx:], 2)
    H = H_f(z) + np.einsum("a,aij->ij", lam_d, T)
    H = 0.5 * (H + H.T)
    q = -(lam_h @ C) if C.shape[0] else np.zeros(len(z))         # pocp.py:357-360
    act = [i for i in range(C.shape[0]) if lam_h[i] != 0]
    return {"A": [S1[:, :nx].copy()], "B": [S1[:, nx:].copy()], "C": [C.copy()], "H": [H], "q": [q],
            "C_As": [C[act] if act else None]}


def dH_of_dP(A, B, dP1, dP2):
    """convexifier.py:181-194: Hessian supplement generated by a storage function change dP (telescoping term)."""
    Q = A.T @ dP2 @ A - dP1
    R = B.T @ dP2 @ B
    Nn = (B.T @ dP2.T @ A).T
    Hc = np.block([[Q, Nn], [Nn.T, R]])
    return 0.5 * (Hc + Hc.T)


def convexify_dare(A, B, H, eps=None, C_As=None, rho=1e-3, scale=None, refine=True):
    """Tuned Hessian Hc = H + dH(dP) > 0 with the same LQ feedback (p = 1).

    dP = P - eps*X, P the stabilising solution of the (indefinite) DARE of (A,B,Q,R,N), X the solution of
    X - Acl' X Acl = I.  In (x, v = u + Kx) coordinates Hc = [[eps I, -eps Acl'XB], [., Rbar - eps B'XB]], positive
    definite for small eps; eps is chosen to (roughly) minimise cond(Hc), the SDP's objective
    (convexifier.py:276,305-306).  With active constraints (type B) rows C_As contribute rho^-1-weighted
    C' C as in convexifier.py:196-200.
    """
    nx = A.shape[0]
    Q, R, Nn = H[:nx, :nx], H[nx:, nx:], H[:nx, nx:]
    if np.min(np.linalg.eigvalsh(H)) > 0:                         # convexifier.py:80-84
        return H.copy(), np.zeros_like(H)
    Hadd = np.zeros_like(H)
    if C_As is not None and len(C_As):
        Hadd = C_As.T @ C_As / rho
    Hq = H + Hadd
    Q, R, Nn = Hq[:nx, :nx], Hq[nx:, nx:], Hq[:nx, nx:]
    P = sla.solve_discrete_are(A, B, Q, R, s=Nn)
    Rbar = R + B.T @ P @ B
    if np.min(np.linalg.eigvalsh(0.5 * (Rbar + Rbar.T))) <= 0:
        raise ValueError("Convexification is not possible: R + B'PB is not positive definite")
    K = np.linalg.solve(Rbar, B.T @ P @ A + Nn.T)
    Acl = A - B @ K
    if np.max(np.abs(np.linalg.eigvals(Acl))) >= 1:
        raise ValueError("Convexification is not possible: LQ closed loop not stable")
    # variable scaling (the reference auto-scales before the SDP, convexifier.py:374-401): weight the Lyapunov
    # equation and measure the condition number in coordinates z / scale
    D = np.ones(H.shape[0]) if scale is None else np.maximum(np.abs(np.asarray(scale, dtype=np.float64)), 1.0)
    X = sla.solve_discrete_lyapunov(Acl.T, np.diag(1.0 / D[:nx] ** 2))

    def build(e):
        return H + Hadd + dH_of_dP(A, B, P - e * X, P - e * X)

    if eps is None:
        best = None
        for e in np.logspace(-8, 4, 400):
            Hc = build(e)
            ev = np.linalg.eigvalsh(Hc * np.outer(D, D))
            if ev[0] > 0 and np.min(np.linalg.eigvalsh(Hc)) > 0:
                cnd = ev[-1] / ev[0]
                if best is None or cnd < best[0]:
                    best = (cnd, e)
        if best is None:
            raise ValueError("Convexification failed")
        eps = best[1]
    Hc = build(eps)
    if refine:
        # the reference's SDP minimises the condition number of the (scaled) tuned Hessian over dP
        # (convexifier.py:276,305-306); same objective, derivative-free search started at the DARE construction
        iu = np.triu_indices(nx)
        Dx = np.outer(D[:nx], D[:nx])
        DD = np.outer(D, D)

        def unpack(v):
            M = np.zeros((nx, nx))
            M[iu] = v
            return (M + M.T - np.diag(np.diag(M))) / Dx

        def obj(v):
            dP = unpack(v)
            ev = np.linalg.eigvalsh((H + Hadd + dH_of_dP(A, B, dP, dP)) * DD)
            return 1e12 * (1.0 - ev[0]) if ev[0] <= 0 else float(np.log(ev[-1] / ev[0]))

        v = ((P - eps * X) * Dx)[iu]
        for _ in range(4):
            v = sopt.minimize(obj, v, method="Nelder-Mead",
                              options={"maxiter": 4000, "xatol": 1e-10, "fatol": 1e-10, "adaptive": True}).x
        dP = unpack(v)
        Hr = H + Hadd + dH_of_dP(A, B, dP, dP)
        if np.min(np.linalg.eigvalsh(Hr)) > 0:
            Hc = Hr
    return Hc, Hc - H


# ------------------------------------------------------------------------------------------------------------------
#  periodic references (p > 1): OCP, sensitivities, convexification
# ------------------------------------------------------------------------------------------------------------------
def solve_periodic_ocp(stage_F, cost_funs, w_guess, nx, alpha_phase=0.1, tol=1e-11, max_iter=200, C=None, c=None):
    """p-periodic OCP (pocp.py:78-203,205-259), optionally with affine path constraints C z_k + c >= 0 at every stage:

        min sum_k l(x_k,u_k) + alpha/2 |x_0 - x0*|^2    s.t.  F(x_k,u_k) - x_{(k+1) mod p} = 0,   C z_k + c >= 0

    With path constraints the pre-solve carries them as inequalities, the rows active at its solution are held as equalities in
    the Newton polish (the active-set re-solve of pocp.py:255-257), and a third return value lam_h (p, nh) holds their
    multipliers (CasADi sign: active => negative).

    The reference first solves with alpha = 0 (IPOPT), then re-solves with the phase-fixing term alpha = 0.1 centred at
    the first solution's x_0 (pocp.py:233-252) -- at that point the term has zero value and zero gradient, so the
    solution is unchanged but the phase-shift direction is no longer (nearly) singular.  Here: SLSQP for alpha = 0,
    then Newton on the phase-fixed KKT system.
    Returns z (p,nz), lam_dyn (p,nx) [CasADi sign: grad f + J' lam = 0], x0star."""
    l_f, g_f, H_f = cost_funs
    z = np.array(w_guess, dtype=np.float64, copy=True)
    p, nz = z.shape
    lam = np.zeros((p, nx))
    nh = 0 if C is None else int(np.asarray(C).shape[0])
    if nh:
        C = np.asarray(C, dtype=np.float64).reshape(nh, nz)
        c = np.asarray(c, dtype=np.float64).ravel()

    def kkt(z, lam, alpha, x0s):
        xf, S, T = stage_F(z[:, :nx], z[:, nx:], 2)
        n, m = p * nz, p * nx
        Hm = np.zeros((n, n))
        J = np.zeros((m, n))
        grad = np.zeros(n)
        r = np.zeros(m)
        for k in range(p):
            sl = slice(k * nz, (k + 1) * nz)
            Hk = H_f(z[k]) + np.einsum("a,aij->ij", lam[k], T[k])
            Hm[sl, sl] = 0.5 * (Hk + Hk.T)
            grad[sl] = g_f(z[k])
            J[k * nx:(k + 1) * nx, sl] = S[k]
            kn = (k + 1) % p
            J[k * nx:(k + 1) * nx, kn * nz:kn * nz + nx] -= np.eye(nx)
            r[k * nx:(k + 1) * nx] = xf[k] - z[kn, :nx]
        if alpha:
            Hm[:nx, :nx] += alpha * np.eye(nx)
            grad[:nx] += alpha * (z[0, :nx] - x0s)
        return Hm, J, grad + J.T @ lam.ravel(), r

    # globalised pre-solve (the reference uses IPOPT here, pocp.py:222-231).  The full-space problem has spurious
    # solutions far from the guess (the unicycle can stand still at the origin with zero cost), so the pre-solve is a
    # feasible-path one: single shooting in (x_0, u_0..u_{p-1}) with the periodicity constraint x_p = x_0, SLSQP with
    # adjoint gradients, phase fix re-centred until x_0 stops moving.  Newton on the phase-fixed KKT system of the
    # multiple-shooting NLP then polishes to round-off and supplies the multipliers.
    nu = nz - nx

    def rollout(v):
        U = v[nx:].reshape(p, nu)
        X = np.zeros((p + 1, nx))
        X[0] = v[:nx]
        A, Bm = [], []
        for k in range(p):
            xf, S1 = stage_F(X[k][None, :], U[k][None, :], 1)
            X[k + 1] = xf[0]
            A.append(S1[0][:, :nx])
            Bm.append(S1[0][:, nx:])
        return X, U, A, Bm

    def ss_fun(v, a, xc):
        X, U, A, Bm = rollout(v)
        Z = np.hstack([X[:p], U])
        f = sum(l_f(zk) for zk in Z) + 0.5 * a * np.sum((v[:nx] - xc) ** 2)
        G = np.array([g_f(zk) for zk in Z])
        adj = np.zeros(nx)
        gu = np.zeros((p, nu))
        for k in range(p - 1, -1, -1):
            gu[k] = G[k, nx:] + Bm[k].T @ adj
            adj = G[k, :nx] + A[k].T @ adj
        return f, np.concatenate([adj + a * (v[:nx] - xc), gu.ravel()])

    def ss_con(v):
        X = rollout(v)[0]
        return X[p] - X[0]

    def ss_jac(v):
        X, U, A, Bm = rollout(v)
        Jc = np.zeros((nx, nx + p * nu))
        Phi = np.eye(nx)
        for k in range(p - 1, -1, -1):
            Jc[:, nx + k * nu:nx + (k + 1) * nu] = Phi @ Bm[k]
            Phi = Phi @ A[k]
        Jc[:, :nx] = Phi - np.eye(nx)
        return Jc

    def ss_h(v):                                        # path constraints along the rollout, (p*nh,)
        X, U, _, _ = rollout(v)
        return (np.hstack([X[:p], U]) @ C.T + c).ravel()

    def ss_hjac(v):
        X, U, A, Bm = rollout(v)
        nv = nx + p * nu
        dX = np.zeros((nx, nv))
        dX[:, :nx] = np.eye(nx)
        Jh = np.zeros((p * nh, nv))
        for k in range(p):
            Jh[k * nh:(k + 1) * nh] = C[:, :nx] @ dX
            Jh[k * nh:(k + 1) * nh, nx + k * nu: nx + (k + 1) * nu] += C[:, nx:]
            dXn = A[k] @ dX
            dXn[:, nx + k * nu: nx + (k + 1) * nu] += Bm[k]
            dX = dXn
        return Jh

    cons = [{"type": "eq", "fun": ss_con, "jac": ss_jac}]
    if nh:
        cons.append({"type": "ineq", "fun": ss_h, "jac": ss_hjac})
    v = np.concatenate([z[0, :nx], z[:, nx:].ravel()])
    for rnd in range(8):
        xc = v[:nx].copy()
        v = sopt.minimize(ss_fun, v, args=(alpha_phase, xc), jac=True, method="SLSQP", constraints=cons,
                          options={"ftol": 1e-15, "maxiter": 500}).x
        if np.max(np.abs(v[:nx] - xc)) < 1e-12:
            break
    X, U, _, _ = rollout(v)
    z = np.hstack([X[:p], U])
    x0s = z[0, :nx].copy()
    if nh:
        return _polish_periodic_with_rows(stage_F, cost_funs, z, nx, C, c, alpha_phase, x0s, tol, max_iter)
    # multipliers: least squares on stationarity, then Newton with the phase fix centred at the pre-solve's x_0
    Hm, J, gl, r = kkt(z, lam, 0.0, x0s)
    lam = np.linalg.lstsq(J.T, -(gl - J.T @ lam.ravel()), rcond=None)[0].reshape(p, nx)
    for it in range(max_iter):
        Hm, J, gl, r = kkt(z, lam, alpha_phase, x0s)
        res_v = np.concatenate([gl, r])
        if np.linalg.norm(res_v, np.inf) < tol:
            break
        n, m = Hm.shape[0], J.shape[0]
        K = np.block([[Hm, J.T], [J, np.zeros((m, m))]])
        step = np.linalg.solve(K, -res_v)
        z = z + step[:n].reshape(p, nz)
        lam = lam + step[n:].reshape(p, nx)
    else:
        raise RuntimeError("periodic OCP: Newton polish did not converge (residual %.2e)" % np.linalg.norm(res_v, np.inf))
    return z, lam, x0s


def _polish_periodic_with_rows(stage_F, cost_funs, z, nx, C, c, alpha, x0s, tol, max_iter):
    """active-set Newton on the KKT system of the multiple-shooting periodic NLP: the rows active at the pre-solve's solution are
    held as equalities; rows whose multiplier comes out with the wrong sign are released, rows that become violated are added."""
    l_f, g_f, H_f = cost_funs
    p, nz = z.shape
    nh = C.shape[0]
    lam = np.zeros((p, nx))
    act = [(k, i) for k in range(p) for i in range(nh) if (C[i] @ z[k] + c[i]) < 1e-7 * max(1.0, abs(c[i]))]
    for outer in range(10):
        mu = np.zeros(len(act))
        for it in range(max_iter):
            xf, S, T = stage_F(z[:, :nx], z[:, nx:], 2)
            n, m, ma = p * nz, p * nx, len(act)
            Hm = np.zeros((n, n)); J = np.zeros((m, n)); Ja = np.zeros((ma, n)); grad = np.zeros(n); r = np.zeros(m); ra = np.zeros(ma)
            for k in range(p):
                sl = slice(k * nz, (k + 1) * nz)
                Hk = H_f(z[k]) + np.einsum("a,aij->ij", lam[k], T[k])
                Hm[sl, sl] = 0.5 * (Hk + Hk.T)
                grad[sl] = g_f(z[k])
                J[k * nx:(k + 1) * nx, sl] = S[k]
                kn = (k + 1) % p
                J[k * nx:(k + 1) * nx, kn * nz:kn * nz + nx] -= np.eye(nx)
                r[k * nx:(k + 1) * nx] = xf[k] - z[kn, :nx]
            Hm[:nx, :nx] += alpha * np.eye(nx)
            grad[:nx] += alpha * (z[0, :nx] - x0s)
            for a, (k, i) in enumerate(act):
                Ja[a, k * nz:(k + 1) * nz] = C[i]
                ra[a] = C[i] @ z[k] + c[i]
            if it == 0 and outer == 0:                      # multiplier estimate from stationarity
                le = np.linalg.lstsq(np.vstack([J, Ja]).T, -grad, rcond=None)[0]
                lam, mu = le[:m].reshape(p, nx), le[m:]
            res_v = np.concatenate([grad + J.T @ lam.ravel() + Ja.T @ mu, r, ra])
            if np.linalg.norm(res_v, np.inf) < tol:
                break
            K = np.block([[Hm, J.T, Ja.T], [J, np.zeros((m, m)), np.zeros((m, ma))], [Ja, np.zeros((ma, m)), np.zeros((ma, ma))]])
            step = np.linalg.solve(K, -res_v)
            z = z + step[:n].reshape(p, nz)
            lam = lam + step[n:n + m].reshape(p, nx)
            mu = mu + step[n + m:]
        else:
            raise RuntimeError("periodic OCP with path constraints: Newton polish did not converge (residual %.2e)" % np.linalg.norm(res_v, np.inf))
        hval = z @ C.T + c
        wrong = [a for a in range(len(act)) if mu[a] > 1e-10]
        viol = [(k, i) for k in range(p) for i in range(nh) if hval[k, i] < -1e-9 and (k, i) not in act]
        if not wrong and not viol:
            break
        act = [e for a, e in enumerate(act) if a not in wrong] + viol
    else:
        raise RuntimeError("periodic OCP with path constraints: the active set did not settle")
    lam_h = np.zeros((p, nh))
    for a, (k, i) in enumerate(act):
        lam_h[k, i] = mu[a]
    lam_h[np.abs(lam_h) < 1e-8] = 0.0
    return z, lam, x0s, lam_h


def sensitivities_periodic(stage_F, cost_funs, z, lam_dyn, nx, alpha_phase=0.1, C=None, lam_h=None):
    """S of pocp.py:261-362 for a p-periodic solution: A_k, B_k, H_k = stage blocks of the Lagrangian Hessian (the phase-fixing
    term alpha*I sits in the x_0 block, as in the reference's NLP; affine path constraints add no curvature),
    q_k = -lam_h,k C (pocp.py:357-360), C_As,k = the rows active at phase k."""
    _, _, H_f = cost_funs
    p, nz = z.shape
    xf, S1, T = stage_F(z[:, :nx], z[:, nx:], 2)
    Hs = []
    for k in range(p):
        Hk = H_f(z[k]) + np.einsum("a,aij->ij", lam_dyn[k], T[k])
        Hk = 0.5 * (Hk + Hk.T)
        if k == 0 and p > 1:
            Hk[:nx, :nx] += alpha_phase * np.eye(nx)
        Hs.append(Hk)
    if C is not None and lam_h is not None and np.asarray(C).shape[0]:
        C = np.asarray(C, dtype=np.float64)
        q = [-(lam_h[k] @ C) for k in range(p)]
        C_As = [C[[i for i in range(C.shape[0]) if lam_h[k, i] != 0]] if np.any(lam_h[k] != 0) else None for k in range(p)]
        return {"A": [S1[k][:, :nx].copy() for k in range(p)], "B": [S1[k][:, nx:].copy() for k in range(p)],
                "C": [C.copy() for _ in range(p)], "C_As": C_As, "H": Hs, "q": q}
    return {"A": [S1[k][:, :nx].copy() for k in range(p)], "B": [S1[k][:, nx:].copy() for k in range(p)],
            "C": None, "C_As": None, "H": Hs, "q": [np.zeros(nz) for _ in range(p)]}


def periodic_riccati(A, B, H):
    """periodic stabilising solution P_k of the (indefinite-cost) Riccati difference equation
        P_k = Q_k + A_k'P_{k+1}A_k - (N_k + A_k'P_{k+1}B_k)(R_k + B_k'P_{k+1}B_k)^-1 (.)',   P_p = P_0.
    P_0 from the DARE of the system lifted over one period (x_0 -> x_p, inputs u_0..u_{p-1}, condensed cost), the other
    phases by one backward sweep from P_p = P_0.  Returns P (p,nx,nx), K (p,nu,nx)."""
    p = len(A)
    nx, nu = B[0].shape
    Mx = np.zeros((p * nx, nx))
    Mu = np.zeros((p * nx, p * nu))
    Phi = np.eye(nx)
    Gam = np.zeros((nx, p * nu))
    for k in range(p):
        Mx[k * nx:(k + 1) * nx] = Phi
        Mu[k * nx:(k + 1) * nx] = Gam
        Gam = A[k] @ Gam
        Gam[:, k * nu:(k + 1) * nu] = B[k]
        Phi = A[k] @ Phi
    Qb = np.zeros((p * nx, p * nx))
    Rb = np.zeros((p * nu, p * nu))
    Nb = np.zeros((p * nx, p * nu))
    for k in range(p):
        Qb[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] = H[k][:nx, :nx]
        Rb[k * nu:(k + 1) * nu, k * nu:(k + 1) * nu] = H[k][nx:, nx:]
        Nb[k * nx:(k + 1) * nx, k * nu:(k + 1) * nu] = H[k][:nx, nx:]
    Ql = Mx.T @ Qb @ Mx
    Nl = Mx.T @ (Qb @ Mu + Nb)
    Rl = Mu.T @ Qb @ Mu + Mu.T @ Nb + Nb.T @ Mu + Rb
    P0 = sla.solve_discrete_are(Phi, Gam, 0.5 * (Ql + Ql.T), 0.5 * (Rl + Rl.T), s=Nl)
    P = [None] * p
    K = [None] * p
    Pn = 0.5 * (P0 + P0.T)
    for k in range(p - 1, -1, -1):
        Q, R, Nn = H[k][:nx, :nx], H[k][nx:, nx:], H[k][:nx, nx:]
        Rk = R + B[k].T @ Pn @ B[k]
        if np.min(np.linalg.eigvalsh(0.5 * (Rk + Rk.T))) <= 0:
            raise ValueError("Convexification is not possible: R + B'PB is not positive definite (stage %d)" % k)
        G = Nn.T + B[k].T @ Pn @ A[k]
        K[k] = np.linalg.solve(Rk, G)
        Pk = Q + A[k].T @ Pn @ A[k] - G.T @ K[k]
        P[k] = 0.5 * (Pk + Pk.T)
        Pn = P[k]
    if np.max(np.abs(P[0] - P0)) > 1e-6 * max(1.0, np.max(np.abs(P0))):
        raise ValueError("periodic Riccati: sweep does not close (|P_0 - P_p| = %.2e)" % np.max(np.abs(P[0] - P0)))
    return np.array(P), np.array(K)


def convexify_periodic(A, B, H, scale=None):
    """Periodic tuned Hessians Hc_k = H_k + dH(dP_k, dP_{k+1}) > 0 (convexifier.py:165-211 is the dH map; the SDP that
    picks dP is replaced by dP_k = P_k - eps X_k with P the periodic Riccati solution and X the periodic Lyapunov
    solution of X_k - Acl_k' X_{k+1} Acl_k = D^-2).  Same LQ feedback as the indefinite problem by construction."""
    p = len(A)
    nx = A[0].shape[0]
    if all(np.min(np.linalg.eigvalsh(Hk)) > 0 for Hk in H):       # convexifier.py:80-84
        return [Hk.copy() for Hk in H]
    P, K = periodic_riccati(A, B, H)
    Acl = [A[k] - B[k] @ K[k] for k in range(p)]
    mono = np.eye(nx)
    for k in range(p):
        mono = Acl[k] @ mono
    if np.max(np.abs(np.linalg.eigvals(mono))) >= 1:
        raise ValueError("Convexification is not possible: periodic LQ closed loop not stable")
    D = np.ones(H[0].shape[0]) if scale is None else np.maximum(np.abs(np.asarray(scale, dtype=np.float64)), 1.0)
    W = np.diag(1.0 / D[:nx] ** 2)
    X = [np.zeros((nx, nx)) for _ in range(p)]
    Xn = np.zeros((nx, nx))
    for sweep in range(20000):
        delta = 0.0
        for k in range(p - 1, -1, -1):
            Xk = W + Acl[k].T @ Xn @ Acl[k]
            delta = max(delta, np.max(np.abs(Xk - X[k])) / max(1.0, np.max(np.abs(Xk))))
            X[k] = Xk
            Xn = Xk
        if delta < 1e-13:
            break

    def build(e):
        return [H[k] + dH_of_dP(A[k], B[k], P[k] - e * X[k], P[(k + 1) % p] - e * X[(k + 1) % p]) for k in range(p)]

    DD = np.outer(D, D)
    best = None
    for e in np.logspace(-8, 4, 400):
        Hc = build(e)
        ev = [np.linalg.eigvalsh(Hk * DD) for Hk in Hc]
        lo = min(v[0] for v in ev)
        if lo > 0 and all(np.min(np.linalg.eigvalsh(Hk)) > 0 for Hk in Hc):
            cnd = max(v[-1] for v in ev) / lo
            if best is None or cnd < best[0]:
                best = (cnd, e)
    if best is None:
        raise ValueError("Convexification failed")
    return build(best[1])

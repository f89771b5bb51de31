"""Path-constraint reformulations on the host, in matrix form (the device only ever sees h = C z + c >= 0 over the stage
variables z = (x, u, us, usc) plus the compiled nonlinear rows g = h_nl(x,u) - us).

What the reference does with CasADi Functions in `tunempc/preprocessing.py`:
  * `input_formatting` / `detect_nonlinear_inequalities` (:35-118): rows of the user's h(x,u) >= 0 that are not affine get a slack
    each, h_nl,i(x,u) - us_i = 0, and the row us_i >= 0 is appended after the affine rows;
  * `add_mpc_slacks` (:120-155): rows of h (all of them, or the ones active somewhere along the reference) are softened,
    h_i + usc_j >= 0, the rows usc >= 0 are appended and every slack is charged 1e3 * max_k(-lam_h[k, i]).
Here the same two steps act on the affine data (C, c) and on sympy expressions of the model card -- nothing symbolic survives
into the controller: the nonlinear rows become `OdeModel.gnl` (code-generated next to the ODE), everything else is a matrix.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import numpy as np
import sympy as sp


def split_path_constraints(x: Sequence[sp.Symbol], u: Sequence[sp.Symbol], h_exprs: Sequence[sp.Expr]):
    """Separate the user's rows h_i(x,u) >= 0 into affine rows and rows that need a slack.

    A row is affine iff its Hessian w.r.t. (x,u) vanishes identically (the reference asks CasADi `which_depends(expr, vars, 2)`,
    preprocessing.py:97).  Returns (C, c, gnl, where): C (n_aff + ns, nx+nu+ns), c -- the affine rows in their original
    order followed by the rows us >= 0 (preprocessing.py:110-112); gnl -- the nonlinear expressions in their original order
    (their slack is column nx+nu+j); where[i] = ('h', row) or ('g', j) tells where user row i went."""
    z = list(x) + list(u)
    nzm = len(z)
    zero = {s: 0 for s in z}
    aff, gnl, where = [], [], []
    for e in h_exprs:
        e = sp.sympify(e)
        curved = any(sp.simplify(sp.diff(e, a, b)) != 0 for i, a in enumerate(z) for b in z[i:])
        if curved:
            where.append(("g", len(gnl)))
            gnl.append(e)
        else:
            where.append(("h", len(aff)))
            aff.append(([float(sp.diff(e, s)) for s in z], float(e.xreplace(zero))))
    ns = len(gnl)
    C = np.zeros((len(aff) + ns, nzm + ns))
    c = np.zeros(len(aff) + ns)
    for r, (row, off) in enumerate(aff):
        C[r, :nzm] = row
        c[r] = off
    for j in range(ns):
        C[len(aff) + j, nzm + j] = 1.0                       # us_j >= 0
    return C, c, gnl, where


def state_only_rows(x: Sequence[sp.Symbol], u: Sequence[sp.Symbol], gnl: Sequence[sp.Expr]) -> List[int]:
    """nonlinear rows that do not depend on the inputs (pmpc.py:1107-1114: their slack row is dropped at stage 0)"""
    us = set(u)
    return [j for j, e in enumerate(gnl) if not (set(sp.sympify(e).free_symbols) & us)]


def soften_rows(C: np.ndarray, c: np.ndarray, lam_h: np.ndarray, slack_flag: str = "active") -> Tuple[np.ndarray, np.ndarray, np.ndarray, List[int]]:
    """Soft constraints for the MPC (preprocessing.py:120-155).  C (nh, nz), c (nh,), lam_h (p, nh) multipliers of h along the
    reference (CasADi sign: active => negative).  slack_flag 'all': every row, 'active': rows with a non-zero multiplier
    at some phase (the reference's `indeces_As`), 'none': nothing.
    Returns (C_soft (nh+nsc, nz+nsc), c_soft, scost (nsc,), rows): h_i + usc_j >= 0 on the chosen rows, the rows usc >= 0
    appended, scost_j = 1e3 * max_k(-lam_h[k, i]) (:145)."""
    if slack_flag not in ("none", "all", "active"):
        raise ValueError("slack_flag must be 'none', 'all' or 'active'")
    nh, nz = C.shape
    lam_h = np.atleast_2d(np.asarray(lam_h, dtype=np.float64)).reshape(-1, nh) if nh else np.zeros((1, 0))
    if slack_flag == "none" or nh == 0:
        return C.copy(), c.copy(), np.zeros(0), []
    rows = [i for i in range(nh) if slack_flag == "all" or np.any(lam_h[:, i] != 0.0)]
    nsc = len(rows)
    if nsc == 0:
        return C.copy(), c.copy(), np.zeros(0), []
    Cs = np.zeros((nh + nsc, nz + nsc))
    Cs[:nh, :nz] = C
    cs = np.concatenate([c, np.zeros(nsc)])
    for j, i in enumerate(rows):
        Cs[i, nz + j] = 1.0                                  # h_i + usc_j
        Cs[nh + j, nz + j] = 1.0                             # usc_j >= 0
    scost = np.array([1e3 * np.max(-lam_h[:, i]) for i in rows])
    return Cs, cs, scost, rows


def soft_model(model, nsc: int):
    """the model card compiled for nsc soft-constraint slacks (the stage width is a compile-time constant of the library)"""
    if nsc == getattr(model, "nsc", 0):
        return model
    base = model.name.split("_sc")[0]
    return dataclasses.replace(model, name=base if nsc == 0 else "%s_sc%d" % (base, nsc), nsc=int(nsc), hess_nz=[])

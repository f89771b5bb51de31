"""Import-guarded CasADi front-end: the user's CasADi Functions -> the model card the rest of the package consumes.

The reference takes `f`, `l`, `h` as CasADi Functions (tunempc/tuner.py:41-88) and hands them to CasADi's own code
generator on its acados path (external/acados/interfaces/acados_template/acados_template/generate_c_code_explicit_ode.py:66-96,
ABI external/acados/acados/utils/external_function_generic.c:199-236).  Here the SX expression graph of each Function
is walked once and rebuilt as sympy expressions, so that `modelgen.generate_header` emits the same `__host__ __device__`
bundle (ODE, Jacobian, structurally non-zero second derivatives, bilinear forms, sparsity tables) as for the built-in
model cards -- the kernels see no difference.

CasADi is not installable in the build container of this repository (no network); everything in this module runs only
where `import casadi` works and is exercised by tests/test_casadi_frontend.py, which is skipped otherwise.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from .modelgen import OdeModel

try:                                                   # pragma: no cover - CasADi absent in the build container
    import casadi as ca
except Exception:                                      # noqa: BLE001
    ca = None


def available():
    return ca is not None


def _require():
    if ca is None:
        raise ImportError("casadi is not installed: use the sympy model cards (tunempc_b200.configs) or install casadi")


def _op_table():
    """CasADi operation codes -> sympy constructors (unary / binary)."""
    t1 = {"OP_NEG": lambda a: -a, "OP_EXP": sp.exp, "OP_LOG": sp.log, "OP_SQRT": sp.sqrt, "OP_SQ": lambda a: a ** 2,
          "OP_SIN": sp.sin, "OP_COS": sp.cos, "OP_TAN": sp.tan, "OP_ASIN": sp.asin, "OP_ACOS": sp.acos, "OP_ATAN": sp.atan,
          "OP_SINH": sp.sinh, "OP_COSH": sp.cosh, "OP_TANH": sp.tanh, "OP_INV": lambda a: 1 / a, "OP_TWICE": lambda a: 2 * a,
          "OP_FABS": sp.Abs, "OP_ASSIGN": lambda a: a}
    t2 = {"OP_ADD": lambda a, b: a + b, "OP_SUB": lambda a, b: a - b, "OP_MUL": lambda a, b: a * b,
          "OP_DIV": lambda a, b: a / b, "OP_POW": lambda a, b: a ** b, "OP_CONSTPOW": lambda a, b: a ** b,
          "OP_ATAN2": sp.atan2, "OP_FMIN": sp.Min, "OP_FMAX": sp.Max}
    u = {getattr(ca, k): v for k, v in t1.items() if hasattr(ca, k)}
    b = {getattr(ca, k): v for k, v in t2.items() if hasattr(ca, k)}
    return u, b


def sx_to_sympy(expr, symbols):
    """one scalar casadi.SX expression -> sympy; `symbols`: {casadi symbol name: sympy Symbol}"""
    _require()
    una, bina = _op_table()
    memo = {}

    def walk(e):
        key = e.__hash__()
        if key in memo:
            return memo[key]
        if e.is_symbolic():
            out = symbols[e.name()]
        elif e.is_constant():
            v = float(e)
            out = sp.Integer(int(v)) if v == int(v) and abs(v) < 2 ** 53 else sp.Float(v, 17)
        else:
            op = e.op()
            if e.n_dep() == 1 and op in una:
                out = una[op](walk(e.dep(0)))
            elif e.n_dep() == 2 and op in bina:
                out = bina[op](walk(e.dep(0)), walk(e.dep(1)))
            else:
                raise NotImplementedError("CasADi operation code %d is not handled by the sympy bridge" % op)
        memo[key] = out
        return out

    return walk(expr)


def model_from_casadi(name, f, l=None, rk_steps=1, tf=1.0, integrator="rk4", discrete=False):
    """casadi.Function f(x,u) -> xdot (or the map x+ for `discrete=True`) and optional stage cost l(x,u) -> OdeModel.
    MX Functions are expanded to SX first (Function.expand)."""
    _require()
    if not f.is_a("SXFunction"):
        f = f.expand()
    nx, nu = f.size1_in(0), f.size1_in(1)
    xs = ca.SX.sym("x", nx)
    us = ca.SX.sym("u", nu)
    xdot = f(xs, us)
    x_sp = sp.symbols("x0:%d" % nx)
    u_sp = sp.symbols("u0:%d" % nu)
    names = {xs[i].name(): x_sp[i] for i in range(nx)}
    names.update({us[i].name(): u_sp[i] for i in range(nu)})
    rhs = [sx_to_sympy(xdot[i], names) for i in range(nx)]
    cost = None
    if l is not None:
        if not l.is_a("SXFunction"):
            l = l.expand()
        cost = sx_to_sympy(l(xs, us)[0], names)
    return OdeModel(name, x_sp, u_sp, rhs, rk_steps=rk_steps, tf=tf, integrator=integrator, discrete=discrete, cost=cost)


def linear_constraints_from_casadi(h, nx, nu):
    """casadi.Function h(x,u) >= 0 -> (C, c) with h = C z + c, for an h whose rows are all affine (CasADi does the differentiation);
    `constraints_from_casadi` handles the general case."""
    _require()
    if not h.is_a("SXFunction"):
        h = h.expand()
    xs = ca.SX.sym("x", nx)
    us = ca.SX.sym("u", nu)
    z = ca.vertcat(xs, us)
    hv = h(xs, us)
    J = ca.jacobian(hv, z)
    if ca.which_depends(hv, z, 2, True).count(True):   # second-order dependence: a nonlinear row (preprocessing.py:91-99)
        raise NotImplementedError("h has nonlinear rows: use constraints_from_casadi (slack reformulation, preprocessing.py:78-118)")
    C = np.array(ca.Function("J", [xs, us], [J])(np.zeros(nx), np.zeros(nu)))
    c = np.array(ca.Function("h0", [xs, us], [hv])(np.zeros(nx), np.zeros(nu))).ravel()
    return C.reshape(-1, nx + nu), c


def constraints_from_casadi(h, model):
    """casadi.Function h(x,u) >= 0 -> the slack form the controller takes (tunempc/preprocessing.py:35-118): rows that are affine in
    (x,u) become (C, c); the others are handed to the model card as `gnl` (code-generated next to the ODE, one slack us_i each, the
    rows us >= 0 appended).  Returns (model with gnl, C over (x,u,us), c, where)."""
    _require()
    import dataclasses
    from .constraints import split_path_constraints
    if not h.is_a("SXFunction"):
        h = h.expand()
    nx, nu = model.nx, model.nu
    xs = ca.SX.sym("x", nx)
    us = ca.SX.sym("u", nu)
    names = {xs[i].name(): model.x[i] for i in range(nx)}
    names.update({us[i].name(): model.u[i] for i in range(nu)})
    hv = h(xs, us)
    rows = [sx_to_sympy(hv[i], names) for i in range(hv.shape[0])]
    C, c, gnl, where = split_path_constraints(model.x, model.u, rows)
    if gnl:
        model = dataclasses.replace(model, gnl=tuple(gnl), hess_nz=[])
    return model, C, c, where


def card_from_casadi(name, f, l, h=None, N=20, p=1, w_guess=None, term_idx=None, **integrator):
    """the model-card dict `Tuner(card, p)` / the `sys` of `Pmpc` take (the counterpart of `Tuner(f, l, h, p)`, tunempc/tuner.py:41).
    Nonlinear rows of h end up in `card['model'].gnl` with their slack rows in C (the Tuner's own OCP handles affine rows only;
    the Pmpc constructor takes the card as it is)."""
    model = model_from_casadi(name, f, l, **integrator)
    nz = model.nx + model.nu
    if h is None:
        C, c = np.zeros((0, nz)), np.zeros(0)
    else:
        model, C, c, _ = constraints_from_casadi(h, model)
    return dict(model=model, cost=model.cost, C=C, c=c, w_guess=np.zeros(nz) if w_guess is None else np.asarray(w_guess, dtype=np.float64),
                period=p, N=N, term_idx=list(range(model.nx)) if term_idx is None else list(term_idx))

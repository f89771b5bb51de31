"""TEST INFRASTRUCTURE (oracle/): sympy restatement of `tunempc/preprocessing.py`, pinned by the known answers of the reference's own
tests (test/test_processing.py).  Nothing in the product package imports it (the product-side counterpart, in matrix form, is tunempc_b200/constraints.py).

  input_formatting(sys)        preprocessing.py:35-76    split h(x,u) >= 0 into linear rows and slacked nonlinear equalities
  detect_nonlinear_inequalities  :78-118                 g(x,u,us) = h_nl(x,u) - us = 0,  h(x,u,us) = [h_lin(x,u); us] >= 0
  add_mpc_slacks(...)          preprocessing.py:120-155  soft constraints: h_i + usc_j >= 0, usc >= 0, L1 penalty scost

The reference works on CasADi Functions and detects (non-)linearity with `ca.which_depends(expr, vars, 2)`; here the
functions are sympy expressions over the model card's symbols and a row is nonlinear iff one of its second derivatives
w.r.t. (x,u) is not identically zero.  Known answers of the reference's own tests (test/test_processing.py:98-104,
177-184) are reproduced in tests/test_preprocessing.py.
"""
from __future__ import annotations

import collections
import itertools

import numpy as np
import sympy as sp


class SymFunction:
    """minimal stand-in for ca.Function: named inputs (tuples of symbols), one vector output"""

    def __init__(self, name, args, exprs):
        self.name = name
        self.args = [tuple(a) for a in args]
        self.exprs = [sp.sympify(e) for e in exprs]
        self._f = sp.lambdify([list(itertools.chain(*self.args))], self.exprs, "numpy")

    def size1_in(self, i):
        return len(self.args[i])

    def size1_out(self, i=0):
        return len(self.exprs)

    def __call__(self, *vals):
        flat = list(itertools.chain(*[np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel().tolist() for v in vals]))
        return np.array(self._f(flat), dtype=np.float64).reshape(-1, 1)

    def subs_exprs(self, *new_args):
        m = {}
        for old, new in zip(self.args, new_args):
            m.update(dict(zip(old, new)))
        return [e.xreplace(m) for e in self.exprs]


def input_formatting(sys):                                              # preprocessing.py:35-76
    fsize = sys["f"][0] if type(sys["f"]) == list else sys["f"]
    nx, nu = fsize.size1_in(0), fsize.size1_in(1)
    sys["vars"] = collections.OrderedDict()
    sys["vars"]["x"] = sp.symbols("x0:%d" % nx)
    sys["vars"]["u"] = sp.symbols("u0:%d" % nu)
    if "h" in sys:
        sys["g"], sys["h"] = detect_nonlinear_inequalities(sys["h"])
        if sys["g"].count(None) == len(sys["g"]):
            del sys["g"]
        else:
            ns = 0
            for k in range(len(sys["g"])):
                ns = sys["g"][k].size1_in(2) if sys["g"][k] is not None else 0
            sys["vars"]["us"] = sp.symbols("us0:%d" % ns)
        if len(sys["h"]) == 1:
            sys["h"] = sys["h"][0]
            if "g" in sys:
                sys["g"] = sys["g"][0]
    return sys


def detect_nonlinear_inequalities(h):                                    # preprocessing.py:78-118
    if type(h) is not list:
        h = [h]
    h_new, g_new = [], []
    for k in range(len(h)):
        x = sp.symbols("x0:%d" % h[k].size1_in(0))
        u = sp.symbols("u0:%d" % h[k].size1_in(1))
        h_expr = h[k].subs_exprs(x, u)
        z = list(x) + list(u)
        h_nlin, h_lin = [], []
        for e in h_expr:                                                 # ca.which_depends(expr, vars, 2)
            second = any(sp.simplify(sp.diff(e, a, b)) != 0 for a in z for b in z)
            (h_nlin if second else h_lin).append(e)
        if len(h_nlin) > 0:
            s = sp.symbols("us0:%d" % len(h_nlin))
            g_new.append(SymFunction("g", [x, u, s], [e - si for e, si in zip(h_nlin, s)]))
            h_new.append(SymFunction("h", [x, u, s], h_lin + list(s)))   # slacks >= 0
        else:
            g_new.append(None)
            h_new.append(h[k])
    return g_new, h_new


def add_mpc_slacks(sys, lam_g, active_set, slack_flag="active"):         # preprocessing.py:120-155
    """lam_g: {'h': array (N, nh)} multipliers of h along the reference (CasADi sign: active => negative)."""
    if ("h" not in sys) or (slack_flag == "none"):
        return sys
    active_constraints = set(itertools.chain(*active_set))
    slack_condition = lambda i: (slack_flag == "all") or ((slack_flag == "active") and (i in active_constraints))
    h_args = list(sys["vars"].values())
    h_expr = sys["h"].subs_exprs(*h_args)
    slacks = [slack_condition(i) for i in range(len(h_expr))]
    if sum(slacks) > 0:
        usc = sp.symbols("usc0:%d" % sum(slacks))
        slack_cost, h_slack = [], []
        j = 0
        for i in range(len(h_expr)):
            if slacks[i]:
                h_slack.append(usc[j])
                slack_cost.append(1e3 * np.max(-np.asarray(lam_g["h"], dtype=np.float64)[:, i]))   # :145
                j += 1
            else:
                h_slack.append(0.0)
        h_new = [e + s for e, s in zip(h_expr, h_slack)] + list(usc)
        sys["h"] = SymFunction("h", h_args + [usc], h_new)
        sys["scost"] = np.array(slack_cost, dtype=np.float64).reshape(-1, 1)
        sys["vars"]["usc"] = usc
    return sys

"""Host-side mirror of the reference's user API `tunempc.Tuner` (`tunempc/tuner.py:39-264`).

    tuner = Tuner(f, l, h, p)          # reference: CasADi Functions f (integrator), l, h ; here: a model card
    wsol  = tuner.solve_ocp(w0)        # tuner.py:90-132   -> pocp.Pocp.solve
    Hc    = tuner.convexify(rho=...)   # tuner.py:134-160  -> convexifier.convexify
    ctrl  = tuner.create_mpc('tuned', N, opts)            # tuner.py:162-199 -> pmpc.Pmpc
    ctrl.step(x0) / ctrl.step(X0)      # the batched CUDA solve (tunempc_b200.pmpc.Pmpc)

CasADi is not available in this image, so `f`, `l`, `h` are given as a *model card*: a dict with the sympy ODE
(`modelgen.OdeModel`), the sympy stage cost and the linear path constraints h(x,u) = C z + c >= 0 (see
`tunempc_b200/configs.py`; every card restates one of the reference's example scripts).  Everything in this module runs
once on the host (offline OCP solve and convexification: out of scope for CUDA); the
controller it returns runs on the GPU through the C ABI and has no CPU fallback.
"""
from __future__ import annotations

import copy
import os

import numpy as np

from . import tuning
from .problem import MpcProblem


class Tuner(object):
    def __init__(self, f, l=None, h=None, p=1, stage_eval=None):
        """f: a model card (dict from `configs.<name>()`) or an `OdeModel`; l: sympy stage cost (default: the card's);
        h: (C, c) of the linear path constraints h = C z + c >= 0 (default: the card's); p: period of the OCP.
        stage_eval(x, u, order): host evaluation of the compiled model's interval map (default: the C-ABI library's
        `tmpc_stage_eval_host`, which needs libtmpc_<model>.so but no GPU)."""
        card = f if isinstance(f, dict) else {"model": f}
        self.__card = card
        self.__model = card["model"]
        self.__l = l if l is not None else card.get("cost")
        if self.__l is None:
            raise ValueError("Tuner needs a stage cost l")
        if h is not None:
            self.__C, self.__c = np.atleast_2d(np.asarray(h[0], dtype=np.float64)), np.asarray(h[1], dtype=np.float64).ravel()
        else:
            self.__C = np.asarray(card.get("C", np.zeros((0, self.__model.nx + self.__model.nu))), dtype=np.float64)
            self.__c = np.asarray(card.get("c", np.zeros(0)), dtype=np.float64)
        self.__nx, self.__nu = self.__model.nx, self.__model.nu
        self.__ns = getattr(self.__model, "ns", 0)                          # slacks of the model card's nonlinear rows (tuner.py:62-66)
        self.__nw = self.__nx + self.__nu + self.__ns                       # tuner.py:69
        if self.__ns:
            if self.__C.shape[1] != self.__nw or self.__C.shape[0] < self.__ns:
                raise ValueError("with nonlinear rows the path constraints are C (x,u,us) + c >= 0, the rows us >= 0 last "
                                 "(constraints.split_path_constraints)")
            if int(p) != 1:
                raise NotImplementedError("nonlinear path constraints in a periodic OCP are not built")
            self.__gnl_funs = tuning.lambdify_gnl(self.__model)
        self.__p = int(p)
        self.__cost_funs = tuning.lambdify_cost(self.__model, self.__l)
        if stage_eval is None:
            from .lib import ModelLib
            stage_eval = ModelLib(self.__model.name).stage_eval
        self.__F = stage_eval
        self.__w_sol = None
        self.__S = None
        self.__lam_h = None
        self.__lam_dyn = None

    # ---- tuner.py:90-132 ---------------------------------------------------------------------------------
    def solve_ocp(self, w0=None, lam0=None):
        """p-periodic OCP (steady state for p = 1).  w0: flat (p*(nx+nu),) or (p, nx+nu) initial guess.  Returns the
        solution as (p, nx+nu)."""
        if w0 is None:
            w0 = self.__card.get("w_guess")
        w0 = np.asarray(w0, dtype=np.float64)
        if self.__ns and w0.size == self.__p * (self.__nx + self.__nu):       # guess without slacks: us = h_nl(x,u)
            w0 = np.concatenate([w0.ravel(), self.__gnl_funs[0](w0.ravel())])
        w0_shape = self.__p * self.__nw
        assert w0.size == w0_shape, \
            "Incorrect dimensions of input variable w0: expected {}x1, but received {}".format(w0_shape, w0.shape)
        w0 = w0.reshape(self.__p, self.__nw)
        self.__lam_gnl = None
        if self.__p == 1 and self.__ns:
            nzm = self.__nx + self.__nu
            z, lam_d, lam_g, lam_h = tuning.solve_steady_state_slack(self.__F, self.__cost_funs, self.__gnl_funs, self.__C, self.__c,
                                                                     w0[0][:nzm], self.__nx, self.__ns)
            self.__S = tuning.sensitivities_slack(self.__F, self.__cost_funs, self.__gnl_funs, self.__C, z, lam_d, lam_g, lam_h,
                                                  self.__nx, self.__nu, self.__ns)
            self.__w_sol = z[None, :].copy()
            self.__lam_h, self.__lam_dyn, self.__lam_gnl = lam_h[None, :], lam_d[None, :], lam_g[None, :]
        elif self.__p == 1:
            z, lam_d, lam_h = tuning.solve_steady_state(self.__F, self.__cost_funs, self.__C, self.__c, w0[0], self.__nx)
            self.__S = tuning.sensitivities(self.__F, self.__cost_funs, self.__C, z, lam_d, lam_h, self.__nx)
            self.__w_sol = z[None, :].copy()
            self.__lam_h, self.__lam_dyn = lam_h[None, :], lam_d[None, :]
        else:
            if self.__C.shape[0]:                                           # pocp.py:205-259 with path constraints
                z, lam_d, _, lam_h = tuning.solve_periodic_ocp(self.__F, self.__cost_funs, w0, self.__nx, C=self.__C, c=self.__c)
                self.__S = tuning.sensitivities_periodic(self.__F, self.__cost_funs, z, lam_d, self.__nx, C=self.__C, lam_h=lam_h)
            else:
                z, lam_d, _ = tuning.solve_periodic_ocp(self.__F, self.__cost_funs, w0, self.__nx)
                lam_h = np.zeros((self.__p, 0))
                self.__S = tuning.sensitivities_periodic(self.__F, self.__cost_funs, z, lam_d, self.__nx)
            self.__w_sol = z.copy()
            self.__lam_h, self.__lam_dyn = lam_h, lam_d
        return self.__w_sol

    # ---- tuner.py:134-160 --------------------------------------------------------------------------------
    def convexify(self, rho=1.0, force=False, solver="dare"):
        """Positive definite stage cost matrices of a tracking NMPC scheme that is locally first-order equivalent to
        economic MPC.  `solver` names the SDP back-end in the reference ('cvxopt' / 'mosek'); here the convexifying
        storage-function change dP comes from a (periodic) Riccati + Lyapunov construction (tuning.py), any value is
        accepted.  As in the reference (convexifier.py:80-84) an already positive definite H is returned unchanged;
        `force` is accepted for signature compatibility and ignored."""
        if self.__S is None:
            raise RuntimeError("call solve_ocp() first")
        S = self.__S
        if self.__p == 1:
            Hc = [tuning.convexify_dare(S["A"][0], S["B"][0], S["H"][0], C_As=S["C_As"][0], rho=rho, scale=self.__w_sol[0])[0]]
        else:
            try:
                Hc = tuning.convexify_periodic(S["A"], S["B"], S["H"])
            except np.linalg.LinAlgError as e:
                # the periodic Riccati construction works on the unconstrained LQ problem along the trajectory; with path
                # constraints active on the orbit that problem need not have a stabilising solution (the reference's SDP
                # convexifies on the null space of the active rows, convexifier.py:213-308 -- not built for p > 1)
                raise ValueError("Convexification failed along the periodic trajectory (%s); with active path constraints use "
                                 "create_mpc('economic') or create_mpc('tracking', tuning=...)" % e)
        S["Hc"] = Hc
        return S["Hc"]

    # ---- tuner.py:162-199 --------------------------------------------------------------------------------
    def create_mpc(self, mpc_type, N, opts={}, tuning=None, device=0):
        """Create an MPC controller of the given type and horizon.  'tuned': H = Hc, q = S['q']; 'tracking': user
        tuning {'H': [...], 'q': [...]} (tuner.py:191-195); 'economic': the stage cost l itself with the OCP's multipliers as
        dual reference and the exact Hessian (tuner.py:180-182, pmpc.py:97-107), steady-state and periodic references."""
        if mpc_type not in ["economic", "tuned", "tracking"]:
            raise ValueError("Provided MPC type not supported.")                           # tuner.py:168-169
        if self.__w_sol is None:
            raise RuntimeError("call solve_ocp() first")
        if mpc_type == "tracking":
            if tuning is None:
                raise ValueError("Tracking type MPC controller requires user-provided tuning!")   # tuner.py:193
        elif mpc_type == "tuned":
            if "Hc" not in self.__S:
                raise RuntimeError("call convexify() first")
            tuning = {"H": self.__S["Hc"], "q": self.__S["q"]}                              # tuner.py:195
        from .pmpc import Pmpc
        opts = dict(opts)
        if opts.get("p_operator") is None:
            opts["p_operator"] = self.__card.get("term_idx", list(range(self.__nx)))
        p = self.__p
        mpc_sys = self.sys
        if opts.get("slack_flag", "none") != "none":                                         # tuner.py:171-177
            # soft constraints (preprocessing.add_mpc_slacks): rows softened by usc >= 0 with an L1 penalty; the stage width is a
            # compile-time constant of the model library, so the controller runs on the library variant <model>_sc<nsc>
            from . import constraints, modelgen
            from .lib import build_model_lib
            Cs, cs, scost, rows = constraints.soften_rows(self.__C, self.__c, self.__lam_h, opts["slack_flag"])
            if len(rows):
                model = constraints.soft_model(self.__model, len(rows))
                hdr = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "gen", "model_%s.h" % model.name)
                modelgen.generate_header(model, hdr)
                build_model_lib(model.name)                                                  # no-op when the library is up to date
                mpc_sys = dict(self.sys, f=model, h=(Cs, cs), scost=scost)
                mpc_sys["vars"] = dict(self.sys["vars"], usc=list(range(len(rows))))
        nzm = self.__nx + self.__nu
        wref = {"x": [self.__w_sol[k, :self.__nx] for k in range(p)], "u": [self.__w_sol[k, self.__nx:nzm] for k in range(p)]}
        if self.__ns:
            wref["us"] = [self.__w_sol[k, nzm:] for k in range(p)]
        sens = {"A": self.__S["A"], "B": [np.asarray(b)[:, :self.__nu] for b in self.__S["B"]]}
        if mpc_type == "economic":                                                           # tuner.py:180-182: full lam_g
            lam_g_ref = {"dyn": list(self.__lam_dyn), "h": list(self.__lam_h)}
            if self.__ns:
                lam_g_ref["g"] = list(self.__lam_gnl)
            return Pmpc(N=N, sys=mpc_sys, cost="economic", wref=wref, lam_g_ref=lam_g_ref, sensitivities=sens,
                        options=opts, device=device)
        lam_g0 = {"dyn": [np.zeros(self.__nx)] * p, "h": list(self.__lam_h)}                # tuner.py:186-189: lam_g0['dyn'] = 0
        if self.__ns:
            lam_g0["g"] = [np.zeros(self.__ns)] * p                                          # tuner.py:188-189: lam_g0['g'] = 0
        return Pmpc(N=N, sys=mpc_sys, cost="tracking", wref=wref, tuning=tuning, lam_g_ref=lam_g0, sensitivities=sens,
                    options=opts, device=device)

    # ---- properties (tuner.py:201-264) -------------------------------------------------------------------
    @property
    def S(self):
        return self.__S

    @property
    def sys(self):
        sys = {"f": self.__model, "h": (self.__C, self.__c), "vars": {"x": self.__model.x, "u": self.__model.u}}
        if self.__ns:
            sys["vars"]["us"] = list(range(self.__ns))
            sys["g"] = "compiled"                                    # the nonlinear rows live in the model library (OdeModel.gnl)
        return sys

    @property
    def l(self):
        return self.__cost_funs[0]

    @property
    def w_sol(self):
        return self.__w_sol

    @property
    def lam_g(self):
        return {"dyn": self.__lam_dyn, "h": self.__lam_h}

    @property
    def p(self):
        return self.__p

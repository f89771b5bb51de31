"""Problem IR for the tuned tracking-MPC feedback solve, and the host-side phase tables.

`MpcProblem` is what `Pmpc.__init__` receives in the reference (`tunempc/pmpc.py:39`: N, sys, cost, wref, tuning,
lam_g_ref, sensitivities, options) minus the CasADi objects: dimensions, the compiled model name, the *linear*
path constraints h = C z + c >= 0, the terminal operator Jacobian T, the periodic reference, the tuning (H, q) and
the reference multipliers.  `build_tables` restates `Pmpc.__create_reference` (`tunempc/pmpc.py:676-783`).

Layouts (`tunempc/pmpc.py:217-256`):
  w   = [x_0,u_0,(us_0),(usc_0), ..., x_{N-1},u_{N-1},(us),(usc), x_N]      n_w = N*nz + nx,  nz = nx+nu+ns+nsc
  g   = [init(nx) | k<N: dyn_k(nx), g_k(ns), h_k(nh) | term(nx_term)]       n_g = nx + N*(nx+ns+nh) + nx_term
  lam = same order as g, CasADi sign convention  grad f + J' lam = 0.
Slack formulation (`tunempc/preprocessing.py:78-155`): the ns nonlinear path constraints h_nl(x,u) >= 0 of the model card
become g_k = h_nl(x_k,u_k) - us_k = 0 with us_k >= 0 among the rows of h; nsc rows of h are softened by usc >= 0 with the
linear cost scost'usc.  h stays linear in the stage variables: h = C z + c with C (nh, nz),
rows = [h_lin (+usc on slacked rows); us; usc].  The reference and the tuning have no usc entries (nzr = nx+nu+ns).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class MpcProblem:
    name: str                      # compiled model (libtmpc_<name>.so / liborc_<name>.so)
    nx: int
    nu: int
    N: int                         # horizon
    p: int                         # period of the reference
    wref: np.ndarray               # (p, nzr)  periodic reference (x_k, u_k, us_k)
    H: np.ndarray                  # (p, nzr, nzr) tuned stage Hessians
    q: np.ndarray                  # (p, nzr)  tuned stage gradients
    C: np.ndarray                  # (nh, nz)  h(x,u,us,usc) = C z + c >= 0
    c: np.ndarray                  # (nh,)
    lam_h_ref: np.ndarray          # (p, nh - nsc) reference multipliers of h (CasADi sign: active => negative); the usc >= 0
                                   #           rows get -scost (pmpc.py:716-720)
    lam_dyn_ref: np.ndarray        # (p, nx)   zero for tuned/tracking controllers (tunempc/tuner.py:186-189)
    term_idx: List[int]            # p_operator = x[term_idx]  (selection; identity by default, pmpc.py:156)
    S_A: Optional[np.ndarray] = None   # (p, nx, nx) dF/dx along the reference (terminal multiplier projection)
    S_B: Optional[np.ndarray] = None   # (p, nx, nu)
    hessian_approximation: str = "exact"   # pmpc.py:153
    max_iter: int = 2000                   # pmpc.py:155
    tol: float = 1e-6                      # sqp_method.py:55
    mpc_type: str = "tuned"                # 'tuned' / 'tracking' (cost from H, q) or 'economic' (cost = the model card's l)
    meta: dict = field(default_factory=dict)
    ns: int = 0                            # slacks of the nonlinear path constraints (the compiled model's gnl rows)
    nsc: int = 0                           # soft-constraint slacks
    scost: Optional[np.ndarray] = None     # (nsc,) linear slack cost (preprocessing.py:145)
    lam_g_ref: Optional[np.ndarray] = None # (p, ns) reference multipliers of the rows g (pmpc.py:713-714)
    gnl_x_idx: List[int] = field(default_factory=list)   # nonlinear constraints that depend on x only (pmpc.py:1107-1114)

    # ---- sizes -----------------------------------------------------------------------------------
    @property
    def nz(self):
        return self.nx + self.nu + self.ns + self.nsc

    @property
    def nzr(self):
        """stage width of the reference / tuning: (x, u, us) -- no usc entries (pmpc.py:186-208)"""
        return self.nx + self.nu + self.ns

    @property
    def nh(self):
        return int(self.C.shape[0])

    @property
    def nx_term(self):
        return len(self.term_idx)

    @property
    def n_w(self):
        return self.N * self.nz + self.nx

    @property
    def gs(self):
        return self.nx + self.ns + self.nh

    @property
    def n_g(self):
        return self.nx + self.N * self.gs + self.nx_term

    @property
    def T(self):
        T = np.zeros((self.nx_term, self.nx))
        for r, i in enumerate(self.term_idx):
            T[r, i] = 1.0
        return T

    @property
    def h_x_idx(self):
        """rows of h that depend on x only -- relaxed at stage 0 (pmpc.py:70,293-294)."""
        return [i for i in range(self.nh) if not np.any(self.C[i, self.nx:] != 0.0)]

    @property
    def h_us_idx(self):
        """pmpc.py:1116, bug-compatible: `idx + nh - ns` for every state-only nonlinear constraint idx.  That is the row
        us_idx >= 0 only when no usc rows follow the us rows; with soft-constraint slacks it points into the usc >= 0 rows
        (SURVEY.md 8(a) quirk 8)."""
        return [int(i) + self.nh - self.ns for i in self.gnl_x_idx]

    @property
    def relax0(self):
        """rows of h whose lower bound is dropped at stage 0 (pmpc.py:293-294: h_us_idx + h_x_idx)"""
        return sorted(set(self.h_us_idx + self.h_x_idx))

    # ---- index helpers ---------------------------------------------------------------------------
    def ix(self, k):
        return slice(k * self.nz, k * self.nz + self.nx)

    def iu(self, k):
        return slice(k * self.nz + self.nx, k * self.nz + self.nx + self.nu)

    def ius(self, k):
        b = k * self.nz + self.nx + self.nu
        return slice(b, b + self.ns)

    def iusc(self, k):
        b = k * self.nz + self.nzr
        return slice(b, b + self.nsc)

    def izr(self, k):
        """(x_k, u_k, us_k): the variables the tracking cost sees (pmpc.py:305-313)"""
        return slice(k * self.nz, k * self.nz + self.nzr)

    def iz(self, k):
        return slice(k * self.nz, (k + 1) * self.nz)

    def g_init(self):
        return slice(0, self.nx)

    def g_dyn(self, k):
        b = self.nx + k * self.gs
        return slice(b, b + self.nx)

    def g_g(self, k):
        b = self.nx + k * self.gs + self.nx
        return slice(b, b + self.ns)

    def g_h(self, k):
        b = self.nx + k * self.gs + self.nx + self.ns
        return slice(b, b + self.nh)

    def g_term(self):
        b = self.nx + self.N * self.gs
        return slice(b, b + self.nx_term)

    def bounds(self):
        """lbg, ubg (pmpc.py:289-294): all 0/0; h in [0, inf); at stage 0 the rows h_us_idx + h_x_idx get lbg = -inf."""
        lbg = np.zeros(self.n_g)
        ubg = np.zeros(self.n_g)
        for k in range(self.N):
            ubg[self.g_h(k)] = np.inf
        if self.nh:
            s0 = self.g_h(0)
            for i in self.relax0:
                lbg[s0.start + i] = -np.inf
        return lbg, ubg

    def device_tables(self):
        """(wref, H, q) padded to the full stage width nz for the C ABI (include/tmpc.h, tmpc_set_tables): zero usc entries in
        wref and H, scost in the usc entries of q -- J += scost'usc (pmpc.py:338-339) is then part of q'(z - wref)."""
        P, nz, nzr = self.p, self.nz, self.nzr
        w = np.zeros((P, nz)); H = np.zeros((P, nz, nz)); q = np.zeros((P, nz))
        w[:, :nzr] = self.wref
        H[:, :nzr, :nzr] = self.H
        q[:, :nzr] = self.q
        if self.nsc:
            q[:, nzr:] = np.asarray(self.scost, dtype=np.float64)[None, :]
        return w, H, q

    # ---- persistence (the '.npz checkpoint' of SURVEY.md section 5) -------------------------------
    def save(self, path):
        np.savez(path, name=self.name, nx=self.nx, nu=self.nu, N=self.N, p=self.p, wref=self.wref, H=self.H,
                 q=self.q, C=self.C, c=self.c, lam_h_ref=self.lam_h_ref, lam_dyn_ref=self.lam_dyn_ref,
                 term_idx=np.array(self.term_idx, dtype=np.int64),
                 S_A=self.S_A if self.S_A is not None else np.zeros(0),
                 S_B=self.S_B if self.S_B is not None else np.zeros(0),
                 hessian_approximation=self.hessian_approximation, max_iter=self.max_iter, tol=self.tol,
                 mpc_type=self.mpc_type, ns=self.ns, nsc=self.nsc,
                 scost=self.scost if self.scost is not None else np.zeros(0),
                 lam_g_ref=self.lam_g_ref if self.lam_g_ref is not None else np.zeros(0),
                 gnl_x_idx=np.array(self.gnl_x_idx, dtype=np.int64))

    @staticmethod
    def load(path):
        d = np.load(path, allow_pickle=False)
        return MpcProblem(name=str(d["name"]), nx=int(d["nx"]), nu=int(d["nu"]), N=int(d["N"]), p=int(d["p"]),
                          wref=d["wref"], H=d["H"], q=d["q"], C=d["C"], c=d["c"], lam_h_ref=d["lam_h_ref"],
                          lam_dyn_ref=d["lam_dyn_ref"], term_idx=[int(i) for i in d["term_idx"]],
                          S_A=d["S_A"] if d["S_A"].size else None, S_B=d["S_B"] if d["S_B"].size else None,
                          hessian_approximation=str(d["hessian_approximation"]), max_iter=int(d["max_iter"]),
                          tol=float(d["tol"]), mpc_type=str(d["mpc_type"]) if "mpc_type" in d else "tuned",
                          ns=int(d["ns"]) if "ns" in d else 0, nsc=int(d["nsc"]) if "nsc" in d else 0,
                          scost=d["scost"] if "scost" in d and d["scost"].size else None,
                          lam_g_ref=d["lam_g_ref"] if "lam_g_ref" in d and d["lam_g_ref"].size else None,
                          gnl_x_idx=[int(i) for i in d["gnl_x_idx"]] if "gnl_x_idx" in d else [])


@dataclass
class Tables:
    ref: np.ndarray      # (p, n_w)        primal reference window per phase (usc entries 0: pmpc.py:930-937)
    ref_du: np.ndarray   # (p, n_g)        dual reference window per phase
    Href: np.ndarray     # (p, N, nzr, nzr)
    qref: np.ndarray     # (p, N, nzr)


def build_tables(pb: MpcProblem) -> Tables:
    """Restatement of Pmpc.__create_reference (tunempc/pmpc.py:676-783)."""
    P, N, nx, nzr = pb.p, pb.N, pb.nx, pb.nzr
    ref = np.zeros((P, pb.n_w))
    ref_du = np.zeros((P, pb.n_g))
    Href = np.zeros((P, N, nzr, nzr))
    qref = np.zeros((P, N, nzr))
    T = pb.T
    for k in range(P):
        for j in range(N):                                           # :694-704 (x, u, us; the usc entries of w0 stay 0, :930-937)
            ref[k, pb.izr(j)] = pb.wref[(k + j) % P]
        ref[k, pb.ix(N)] = pb.wref[(k + N) % P, :nx]                  # :706
        lam = np.zeros(pb.n_g)
        lam[pb.g_init()] = -pb.lam_dyn_ref[(k - 1) % P]               # :710
        for j in range(N):                                           # :711-720
            lam[pb.g_dyn(j)] = pb.lam_dyn_ref[(k + j) % P]
            if pb.ns:
                lam[pb.g_g(j)] = pb.lam_g_ref[(k + j) % P]            # :713-714
            if pb.nh:
                lh = pb.lam_h_ref[(k + j) % P]
                if pb.nsc:
                    lh = np.concatenate([lh, -np.asarray(pb.scost, dtype=np.float64)])   # :716-720 ("TODO not entirely correct")
                lam[pb.g_h(j)] = lh
        lam_last = pb.lam_dyn_ref[(k + N - 1) % P]
        lam[pb.g_term()] = T @ lam_last                               # :721
        if pb.nx_term != nx:                                          # :724-767 terminal multiplier projection
            assert pb.S_A is not None and pb.S_B is not None
            A_m, b_m = [], []
            A_factor = np.eye(nx)
            for j in range(N):
                BtA = pb.S_B[(N - j - 1) % P].T @ A_factor
                A_m.append(BtA @ T.T)
                b_m.append(BtA @ lam_last)
                A_factor = pb.S_A[(N - j - 1) % P].T @ A_factor
            A_m = np.vstack(A_m)
            b_m = np.concatenate(b_m)
            LI, R0 = [], 0
            for i in range(A_m.shape[0]):
                R = np.linalg.matrix_rank(A_m[LI + [i], :])
                if R > R0:
                    LI.append(i)
                    R0 = R
            lam_term = np.linalg.solve(A_m[LI, :], b_m[LI])
            lam[pb.g_term()] = lam_term
            delta = -lam_last + T.T @ lam_term
            sl = pb.g_dyn(N - 1)
            lam[sl] += delta
            for j in range(1, N + 1):
                delta = pb.S_A[(N - j) % P].T @ delta
                if j < N:
                    lam[pb.g_dyn(N - 1 - j)] += delta
                else:
                    lam[pb.g_init()] += -delta
        ref_du[k] = lam
        for j in range(N):                                           # :773-775
            Href[k, j] = pb.H[(k + j) % P]
            qref[k, j] = pb.q[(k + j) % P]
    return Tables(ref, ref_du, Href, qref)

"""Problem IR for the tuned tracking-MPC feedback solve, and the host-side phase tables.

`MpcProblem` is what `Pmpc.__init__` receives in the reference (`tunempc/pmpc.py:39`: N, sys, cost, wref, tuning,
lam_g_ref, sensitivities, options) minus the CasADi objects: dimensions, the compiled model name, the *linear*
path constraints h = C z + c >= 0, the terminal operator Jacobian T, the periodic reference, the tuning (H, q) and
the reference multipliers.  `build_tables` restates `Pmpc.__create_reference` (`tunempc/pmpc.py:676-783`).

Layouts (`tunempc/pmpc.py:217-256`), ns = nsc = 0 in this round:
  w   = [x_0,u_0, x_1,u_1, ..., x_{N-1},u_{N-1}, x_N]                      n_w = N*nz + nx
  g   = [init(nx) | k<N: dyn_k(nx), h_k(nh) | term(nx_term)]                n_g = nx + N*(nx+nh) + nx_term
  lam = same order as g, CasADi sign convention  grad f + J' lam = 0.
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
    wref: np.ndarray               # (p, nz)   periodic reference (x_k, u_k)
    H: np.ndarray                  # (p, nz, nz) tuned stage Hessians
    q: np.ndarray                  # (p, nz)   tuned stage gradients
    C: np.ndarray                  # (nh, nz)  h(x,u) = C z + c >= 0
    c: np.ndarray                  # (nh,)
    lam_h_ref: np.ndarray          # (p, nh)   reference multipliers of h (CasADi sign: active => negative)
    lam_dyn_ref: np.ndarray        # (p, nx)   zero for tuned/tracking controllers (tunempc/tuner.py:186-189)
    term_idx: List[int]            # p_operator = x[term_idx]  (selection; identity by default, pmpc.py:156)
    S_A: Optional[np.ndarray] = None   # (p, nx, nx) dF/dx along the reference (terminal multiplier projection)
    S_B: Optional[np.ndarray] = None   # (p, nx, nu)
    hessian_approximation: str = "exact"   # pmpc.py:153
    max_iter: int = 2000                   # pmpc.py:155
    tol: float = 1e-6                      # sqp_method.py:55
    mpc_type: str = "tuned"                # 'tuned' / 'tracking' (cost from H, q) or 'economic' (cost = the model card's l)
    meta: dict = field(default_factory=dict)

    # ---- sizes -----------------------------------------------------------------------------------
    @property
    def nz(self):
        return self.nx + self.nu

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
    def n_g(self):
        return self.nx + self.N * (self.nx + self.nh) + self.nx_term

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

    # ---- index helpers ---------------------------------------------------------------------------
    def ix(self, k):
        return slice(k * self.nz, k * self.nz + self.nx)

    def iu(self, k):
        return slice(k * self.nz + self.nx, (k + 1) * self.nz)

    def iz(self, k):
        return slice(k * self.nz, (k + 1) * self.nz)

    def g_init(self):
        return slice(0, self.nx)

    def g_dyn(self, k):
        b = self.nx + k * (self.nx + self.nh)
        return slice(b, b + self.nx)

    def g_h(self, k):
        b = self.nx + k * (self.nx + self.nh) + self.nx
        return slice(b, b + self.nh)

    def g_term(self):
        b = self.nx + self.N * (self.nx + self.nh)
        return slice(b, b + self.nx_term)

    def bounds(self):
        """lbg, ubg (pmpc.py:289-294): all 0/0; h in [0, inf); stage-0 state-only rows of h get lbg = -inf."""
        lbg = np.zeros(self.n_g)
        ubg = np.zeros(self.n_g)
        for k in range(self.N):
            ubg[self.g_h(k)] = np.inf
        if self.nh:
            s0 = self.g_h(0)
            for i in self.h_x_idx:
                lbg[s0.start + i] = -np.inf
        return lbg, ubg

    # ---- persistence (the '.npz checkpoint' of SURVEY.md section 5) -------------------------------
    def save(self, path):
        np.savez(path, name=self.name, nx=self.nx, nu=self.nu, N=self.N, p=self.p, wref=self.wref, H=self.H,
                 q=self.q, C=self.C, c=self.c, lam_h_ref=self.lam_h_ref, lam_dyn_ref=self.lam_dyn_ref,
                 term_idx=np.array(self.term_idx, dtype=np.int64),
                 S_A=self.S_A if self.S_A is not None else np.zeros(0),
                 S_B=self.S_B if self.S_B is not None else np.zeros(0),
                 hessian_approximation=self.hessian_approximation, max_iter=self.max_iter, tol=self.tol,
                 mpc_type=self.mpc_type)

    @staticmethod
    def load(path):
        d = np.load(path, allow_pickle=False)
        return MpcProblem(name=str(d["name"]), nx=int(d["nx"]), nu=int(d["nu"]), N=int(d["N"]), p=int(d["p"]),
                          wref=d["wref"], H=d["H"], q=d["q"], C=d["C"], c=d["c"], lam_h_ref=d["lam_h_ref"],
                          lam_dyn_ref=d["lam_dyn_ref"], term_idx=[int(i) for i in d["term_idx"]],
                          S_A=d["S_A"] if d["S_A"].size else None, S_B=d["S_B"] if d["S_B"].size else None,
                          hessian_approximation=str(d["hessian_approximation"]), max_iter=int(d["max_iter"]),
                          tol=float(d["tol"]), mpc_type=str(d["mpc_type"]) if "mpc_type" in d else "tuned")


@dataclass
class Tables:
    ref: np.ndarray      # (p, n_w)        primal reference window per phase
    ref_du: np.ndarray   # (p, n_g)        dual reference window per phase
    Href: np.ndarray     # (p, N, nz, nz)
    qref: np.ndarray     # (p, N, nz)


def build_tables(pb: MpcProblem) -> Tables:
    """Restatement of Pmpc.__create_reference (tunempc/pmpc.py:676-783)."""
    P, N, nx, nz = pb.p, pb.N, pb.nx, pb.nz
    ref = np.zeros((P, pb.n_w))
    ref_du = np.zeros((P, pb.n_g))
    Href = np.zeros((P, N, nz, nz))
    qref = np.zeros((P, N, nz))
    T = pb.T
    for k in range(P):
        for j in range(N):                                           # :694-704
            ref[k, pb.iz(j)] = pb.wref[(k + j) % P]
        ref[k, pb.ix(N)] = pb.wref[(k + N) % P, :nx]                  # :706
        lam = np.zeros(pb.n_g)
        lam[pb.g_init()] = -pb.lam_dyn_ref[(k - 1) % P]               # :710
        for j in range(N):                                           # :711-720
            lam[pb.g_dyn(j)] = pb.lam_dyn_ref[(k + j) % P]
            if pb.nh:
                lam[pb.g_h(j)] = pb.lam_h_ref[(k + j) % P]
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

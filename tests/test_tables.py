"""Independent check of `tunempc_b200.problem.build_tables` (the restatement of Pmpc.__create_reference,
tunempc/pmpc.py:676-783) on a hand-built case: p = 2, N = 3, nx = 2, nu = 1, terminal projection on x[0] (nx_term = 1).
Every expected number below is written out from the reference's formulas by hand (scalars, no shared code path), so a
mistake in the table builder that the oracle would inherit (it imports the same function) shows up here."""
import numpy as np

from tunempc_b200.problem import MpcProblem, build_tables


def _problem():
    nx, nu, N, P = 2, 1, 3, 2
    wref = np.array([[1.0, 2.0, 0.5], [3.0, -1.0, -0.25]])             # (p, nz): x(2) | u(1)
    H = np.stack([np.diag([1.0, 2.0, 3.0]), np.diag([4.0, 5.0, 6.0])])
    q = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]])
    lam_dyn = np.array([[1.5, -0.5], [0.25, 2.0]])
    S_A = np.stack([np.array([[1.0, 0.5], [0.0, 2.0]]), np.array([[0.5, 0.0], [1.0, 1.0]])])
    S_B = np.stack([np.array([[2.0], [1.0]]), np.array([[1.0], [-1.0]])])
    return MpcProblem(name="hand", nx=nx, nu=nu, N=N, p=P, wref=wref, H=H, q=q, C=np.zeros((0, 3)), c=np.zeros(0),
                      lam_h_ref=np.zeros((P, 0)), lam_dyn_ref=lam_dyn, term_idx=[0], S_A=S_A, S_B=S_B)


def test_tables_hand_case():
    pb = _problem()
    tab = build_tables(pb)
    N, P = 3, 2
    # primal window (pmpc.py:692-706): phase k holds wref[(k+j)%p], then x of wref[(k+N)%p]
    assert np.array_equal(tab.ref[0], np.array([1.0, 2.0, 0.5, 3.0, -1.0, -0.25, 1.0, 2.0, 0.5, 3.0, -1.0]))
    assert np.array_equal(tab.ref[1], np.array([3.0, -1.0, -0.25, 1.0, 2.0, 0.5, 3.0, -1.0, -0.25, 1.0, 2.0]))
    # tuning windows (:773-775)
    for k in range(P):
        for j in range(N):
            assert np.array_equal(tab.Href[k, j], pb.H[(k + j) % P]) and np.array_equal(tab.qref[k, j], pb.q[(k + j) % P])
    lam = pb.lam_dyn_ref
    A, B = pb.S_A, pb.S_B
    for k in range(P):
        l_last = lam[(k + N - 1) % P]
        # terminal multiplier (:724-754): rows (B[(N-j-1)%p]' Afac) T' lam_term = (B[(N-j-1)%p]' Afac) lam_last, first
        # row of full rank.  T = [1 0], j = 0: Afac = I, B[(N-1)%p] = B[0] = [2, 1]':  2 lam_term = 2 l0 + 1 l1
        lam_term = (2.0 * l_last[0] + 1.0 * l_last[1]) / 2.0
        delta = np.array([-l_last[0] + lam_term, -l_last[1]])          # :757-760  -lam_last + T' lam_term
        dyn = [lam[(k + j) % P].copy() for j in range(N)]
        dyn[2] = dyn[2] + delta                                       # :761
        d1 = A[(N - 1) % P].T @ delta                                 # j = 1: S_A[(N-1)%p] = S_A[0]   (:762-766)
        dyn[1] = dyn[1] + d1
        d2 = A[(N - 2) % P].T @ d1                                    # j = 2: S_A[1]
        dyn[0] = dyn[0] + d2
        d3 = A[(N - 3) % P].T @ d2                                    # j = 3 = N: S_A[0], goes to 'init' with a minus sign
        init = -lam[(k - 1) % P] - d3                                 # :710, :768-769
        exp = np.concatenate([init, dyn[0], dyn[1], dyn[2], [lam_term]])
        assert np.allclose(tab.ref_du[k], exp, rtol=0, atol=1e-14), (k, tab.ref_du[k], exp)
    # the property the projection exists for: the terminal stationarity  -lam_dyn[N-1] + T' lam_term = 0  holds again
    for k in range(P):
        l_dyn_last = tab.ref_du[k][pb.g_dyn(N - 1)]
        l_term = tab.ref_du[k][pb.g_term()]
        assert np.allclose(-l_dyn_last + pb.T.T @ l_term, 0.0, atol=1e-14)


def test_tables_identity_terminal():
    """nx_term = nx: no projection, 'term' carries lam_dyn[(k+N-1)%p] itself (:721)"""
    pb = _problem()
    pb.term_idx = [0, 1]
    tab = build_tables(pb)
    for k in range(2):
        assert np.array_equal(tab.ref_du[k][pb.g_term()], pb.lam_dyn_ref[(k + 3 - 1) % 2])
        assert np.array_equal(tab.ref_du[k][pb.g_init()], -pb.lam_dyn_ref[(k - 1) % 2])
        for j in range(3):
            assert np.array_equal(tab.ref_du[k][pb.g_dyn(j)], pb.lam_dyn_ref[(k + j) % 2])

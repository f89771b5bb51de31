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


def test_tables_slack_formulation_hand_case():
    """us / usc / g rows (pmpc.py:217-235,242-256,289-294,692-720,930-937): p = 2, N = 2, nx = 1, nu = 1, ns = 1, nsc = 1.
    h before the soft slack = [u + 1 >= 0 ; us >= 0] (preprocessing.py:110-112); row 0 is softened: [u + 1 + usc ; us ; usc]
    (preprocessing.py:140-150).  Layouts, bounds and both tables written out by hand."""
    nx, nu, ns, nsc, N, P = 1, 1, 1, 1, 2, 2
    C = np.array([[0.0, 1.0, 0.0, 1.0],      # u + 1 + usc
                  [0.0, 0.0, 1.0, 0.0],      # us
                  [0.0, 0.0, 0.0, 1.0]])     # usc
    c = np.array([1.0, 0.0, 0.0])
    wref = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])                # (p, nzr): x | u | us -- no usc entry (pmpc.py:186-196)
    H = np.stack([np.diag([1.0, 2.0, 3.0]), np.diag([4.0, 5.0, 6.0])])
    q = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]])
    pb = MpcProblem(name="hand", nx=nx, nu=nu, N=N, p=P, wref=wref, H=H, q=q, C=C, c=c,
                    lam_h_ref=np.array([[-0.5, 0.0], [0.0, -0.25]]), lam_dyn_ref=np.zeros((P, nx)), term_idx=[0],
                    ns=ns, nsc=nsc, scost=np.array([500.0]), lam_g_ref=np.array([[0.7], [-0.9]]), gnl_x_idx=[0])
    assert (pb.nz, pb.nzr, pb.nh, pb.n_w, pb.n_g) == (4, 3, 3, 2 * 4 + 1, 1 + 2 * (1 + 1 + 3) + 1)
    # g order [init | dyn_0 g_0 h_0(3) | dyn_1 g_1 h_1(3) | term]  (pmpc.py:242-256, 279-287)
    assert (pb.g_dyn(0), pb.g_g(0), pb.g_h(0)) == (slice(1, 2), slice(2, 3), slice(3, 6))
    assert (pb.g_dyn(1), pb.g_g(1), pb.g_h(1), pb.g_term()) == (slice(6, 7), slice(7, 8), slice(8, 11), slice(11, 12))
    # stage-0 relaxation (pmpc.py:293-294): no row of h depends on x only; h_us_idx = idx + nh - ns = 0 + 3 - 1 = 2 (pmpc.py:1116)
    # -- with the usc row appended after the us row that is the row `usc >= 0`, not `us >= 0` (bug-compatible)
    assert pb.h_x_idx == [] and pb.h_us_idx == [2] and pb.relax0 == [2]
    lbg, ubg = pb.bounds()
    assert np.array_equal(lbg, np.array([0, 0, 0, 0, 0, -np.inf, 0, 0, 0, 0, 0, 0.0]))
    assert np.array_equal(ubg, np.array([0, 0, 0, np.inf, np.inf, np.inf, 0, 0, np.inf, np.inf, np.inf, 0.0]))
    tab = build_tables(pb)
    # primal window: (x, u, us) of wref[(k+j)%p], usc = 0 (w0['usc'] stays 0, pmpc.py:930-937), then x of wref[(k+N)%p]
    assert np.array_equal(tab.ref[0], np.array([1.0, 2.0, 3.0, 0.0, 4.0, 5.0, 6.0, 0.0, 1.0]))
    assert np.array_equal(tab.ref[1], np.array([4.0, 5.0, 6.0, 0.0, 1.0, 2.0, 3.0, 0.0, 4.0]))
    # dual window (pmpc.py:709-721): g rows lam_g_ref, h rows [lam_h_ref ; -scost]
    assert np.array_equal(tab.ref_du[0], np.array([0, 0, 0.7, -0.5, 0.0, -500.0, 0, -0.9, 0.0, -0.25, -500.0, 0.0]))
    assert np.array_equal(tab.ref_du[1], np.array([0, 0, -0.9, 0.0, -0.25, -500.0, 0, 0.7, -0.5, 0.0, -500.0, 0.0]))
    assert tab.Href.shape == (2, 2, 3, 3) and np.array_equal(tab.Href[1, 1], H[0]) and np.array_equal(tab.qref[1, 0], q[1])
    # tables as the C ABI takes them: nz wide, scost in the usc entry of q (J += scost'usc, pmpc.py:338-339)
    wd, Hd, qd = pb.device_tables()
    assert np.array_equal(wd[1], np.array([4.0, 5.0, 6.0, 0.0])) and np.array_equal(qd[0], np.array([0.1, 0.2, 0.3, 500.0]))
    assert np.array_equal(Hd[0], np.diag([1.0, 2.0, 3.0, 0.0]))


def test_problem_roundtrip_with_slacks(tmp_path):
    from conftest import load_problem
    pb = load_problem("awe9")
    assert (pb.nx, pb.nu, pb.ns, pb.nsc, pb.nh, pb.N, pb.p, pb.nx_term) == (9, 3, 3, 3, 17, 20, 40, 7)
    assert (pb.n_w, pb.n_g) == (369, 596)
    f = str(tmp_path / "pb.npz")
    pb.save(f)
    pb2 = MpcProblem.load(f)
    assert pb2.ns == 3 and pb2.nsc == 3 and np.array_equal(pb2.scost, pb.scost) and pb2.gnl_x_idx == pb.gnl_x_idx
    assert pb2.relax0 == pb.relax0 == sorted(set(pb.h_x_idx + [0 + 17 - 3]))

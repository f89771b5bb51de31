"""Generate the committed fixtures under tests/golden/ (run in the build container; needs oracle/ built).

  problem_<cfg>.npz   the tuned MpcProblem (tables the controller is constructed from)
  golden_<cfg>.npz    seeded x0 batch + the ORACLE's answers (oracle/reference_port.py, QP = the reference tree's
                      qpOASES_e from oracle/_ref): u0, w, lam_g, iter, status, nAS at tol 1e-6 and tol 1e-9

The reference itself cannot run here (CasADi absent), so these are outputs of the restatement, not of the reference:
"parity unpinned" in the sense of SURVEY.md section 8(c).  Independent pins: LQ gain G (dense KKT).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_port as rp            # noqa: E402
from tunempc_b200 import configs                   # noqa: E402
from tunempc_b200.problem import build_tables      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


sample_x0 = configs.sample_x0      # synthetic initial states of SURVEY.md section 8(d) (same generator as bench.py)


def unicycle():
    """config #4: periodic reference (p = N = 30), projected terminal constraint, closed loop with plant = model."""
    name = "unicycle"
    st = rp.StageLib(name)
    pb, info = configs.make_problem(name, st.F)
    pb.save(os.path.join(HERE, "problem_%s.npz" % name))
    B = 16
    X0 = sample_x0(name, pb, B)
    out = {"X0": X0}
    for tag, tol in (("t6", 1e-6), ("t9", 1e-9)):
        ctrl = rp.Pmpc(pb, qp="qpoases", sqp_options={"tol": tol})
        U, W, LAM, IT, ST = [], [], [], [], []
        for b in range(B):
            ctrl.reset()
            u = ctrl.step(X0[b])
            U.append(u); W.append(ctrl.w_sol); LAM.append(ctrl.lam_g)
            IT.append(ctrl.log["iter"][-1]); ST.append(ctrl.log["status"][-1])
        out.update({"u0_" + tag: np.array(U), "w_" + tag: np.array(W), "lam_" + tag: np.array(LAM),
                    "iter_" + tag: np.array(IT), "status_" + tag: np.array(ST)})
        print(name, tag, "iter hist", np.bincount(np.array(IT)), "status", np.bincount(np.array(ST)))
    # closed loop: 8 instances x 6 steps, and 2 instances x 36 steps (wraps the period: phase index 30 -> 0)
    for key, nb, ns in (("cl", 8, 6), ("cll", 2, 36)):
        ctrl = rp.Pmpc(pb, qp="qpoases")
        Xcl, Ucl, Icl = [], [], []
        for b in range(nb):
            ctrl.reset()
            x = X0[b].copy()
            xs_, us_, it_ = [x.copy()], [], []
            for _ in range(ns):
                u = ctrl.step(x)
                x = st.F(x[None, :], u[None, :])[0]
                xs_.append(x.copy()); us_.append(u.copy()); it_.append(ctrl.log["iter"][-1])
                assert ctrl.log["status"][-1] == 0
            Xcl.append(xs_); Ucl.append(us_); Icl.append(it_)
        out[key + "_X"] = np.array(Xcl)
        out[key + "_U"] = np.array(Ucl)
        out[key + "_iter"] = np.array(Icl)
        print(name, key, "iter hist", np.bincount(np.array(Icl).ravel()))
    np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % name), **out)


def evaporation():
    """config #3: collocation integrator, pure state constraints (relaxed at stage 0, pmpc.py:293-294), reference on the
    bound X2 = 25 (non-zero reference multipliers -> type-B tuning with C_As, reduced-space convexification)."""
    name = "evaporation"
    st = rp.StageLib(name)
    pb, info = configs.make_problem(name, st.F)
    pb.save(os.path.join(HERE, "problem_%s.npz" % name))
    B = 24
    X0 = sample_x0(name, pb, B)
    out = {"X0": X0}
    for tag, tol in (("t6", 1e-6), ("t9", 1e-9)):
        ctrl = rp.Pmpc(pb, qp="qpoases", sqp_options={"tol": tol})
        U, W, LAM, IT, ST, NAS = [], [], [], [], [], []
        for b in range(B):
            ctrl.reset()
            u = ctrl.step(X0[b])
            U.append(u); W.append(ctrl.w_sol); LAM.append(ctrl.lam_g)
            IT.append(ctrl.log["iter"][-1]); ST.append(ctrl.log["status"][-1]); NAS.append(ctrl.log["nAS"][-1])
        out.update({"u0_" + tag: np.array(U), "w_" + tag: np.array(W), "lam_" + tag: np.array(LAM),
                    "iter_" + tag: np.array(IT), "status_" + tag: np.array(ST), "nAS_" + tag: np.array(NAS)})
        print(name, tag, "iter hist", np.bincount(np.array(IT)), "status", np.bincount(np.array(ST)), "nAS", np.bincount(np.array(NAS)))
    ctrl = rp.Pmpc(pb, qp="qpoases")
    Xcl, Ucl = [], []
    for b in range(4):
        ctrl.reset()
        x = X0[b].copy()
        xs_, us_ = [x.copy()], []
        for _ in range(5):
            u = ctrl.step(x)
            x = st.F(x[None, :], u[None, :])[0]
            xs_.append(x.copy()); us_.append(u.copy())
        Xcl.append(xs_); Ucl.append(us_)
    out["cl_X"] = np.array(Xcl)
    out["cl_U"] = np.array(Ucl)
    np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % name), **out)


def economic():
    """create_mpc('economic') (tuner.py:180-182; pmpc.py:97-107): stage cost l(x,u), exact Hessian, OCP multipliers as dual
    reference -- the controller the tuned one is first-order equivalent to (closed_loop_tools.check_equivalence)."""
    from tunempc_b200 import tuning
    for name, B in (("cstr", 24), ("evaporation", 16)):
        st = rp.StageLib(name)
        pb, info = configs.make_problem(name, st.F, mpc_type="economic")
        pb.save(os.path.join(HERE, "problem_%s_economic.npz" % name))
        cf = tuning.lambdify_cost(info["cfg"]["model"], info["cfg"]["cost"])
        X0 = sample_x0(name, pb, B)
        out = {"X0": X0}
        for tag, tol in (("t6", 1e-6), ("t9", 1e-9)):
            ctrl = rp.Pmpc(pb, qp="qpoases", sqp_options={"tol": tol}, cost_funs=cf)
            U, W, LAM, IT, ST, NAS = [], [], [], [], [], []
            for b in range(B):
                ctrl.reset()
                u = ctrl.step(X0[b])
                U.append(u); W.append(ctrl.w_sol); LAM.append(ctrl.lam_g)
                IT.append(ctrl.log["iter"][-1]); ST.append(ctrl.log["status"][-1]); NAS.append(ctrl.log["nAS"][-1])
            out.update({"u0_" + tag: np.array(U), "w_" + tag: np.array(W), "lam_" + tag: np.array(LAM),
                        "iter_" + tag: np.array(IT), "status_" + tag: np.array(ST), "nAS_" + tag: np.array(NAS)})
            print(name, "economic", tag, "iter hist", np.bincount(np.array(IT)), "status", np.bincount(np.array(ST)), "n_reg", ctrl.sqp.n_reg)
        np.savez_compressed(os.path.join(HERE, "golden_%s_economic.npz" % name), **out)


def unicycle_economic():
    """economic MPC on the periodic unicycle reference (pmpc.py:97-107 with p = N = 30): 8 instances x 6 closed-loop steps"""
    from tunempc_b200 import tuning
    name = "unicycle"
    st = rp.StageLib(name)
    pb, info = configs.make_problem(name, st.F, mpc_type="economic")
    pb.save(os.path.join(HERE, "problem_unicycle_economic.npz"))
    cf = tuning.lambdify_cost(info["cfg"]["model"], info["cfg"]["cost"])
    B, ns = 8, 6
    X0 = sample_x0(name, pb, B, 3)
    ctrl = rp.Pmpc(pb, qp="qpoases", cost_funs=cf)
    Xcl, Ucl, Icl = [], [], []
    for b in range(B):
        ctrl.reset()
        x = X0[b].copy()
        xs_, us_, it_ = [x.copy()], [], []
        for _ in range(ns):
            u = ctrl.step(x)
            assert ctrl.log["status"][-1] == 0
            x = st.F(x[None, :], u[None, :])[0]
            xs_.append(x.copy()); us_.append(u.copy()); it_.append(ctrl.log["iter"][-1])
        Xcl.append(xs_); Ucl.append(us_); Icl.append(it_)
    np.savez_compressed(os.path.join(HERE, "golden_unicycle_economic.npz"), X0=X0, cl_X=np.array(Xcl), cl_U=np.array(Ucl), cl_iter=np.array(Icl))
    print("unicycle economic: iter hist", np.bincount(np.array(Icl).ravel()))


def _awe9_single(args):
    i, x0, tol = args
    pb = rp_problem("awe9")
    ctrl = rp.Pmpc(pb, qp="qpoases", sqp_options={"tol": tol})
    u = ctrl.step(x0)
    lg = ctrl.log
    return (i, u, ctrl.w_sol, ctrl.lam_g, lg["iter"][-1], lg["status"][-1], lg["f"][-1], lg["nAS"][-1], lg["nACtot"][-1], lg["nAC"][-1])


def _awe9_loop(args):
    b, x0, nsteps = args
    pb = rp_problem("awe9")
    st = rp.StageLib("awe9")
    ctrl = rp.Pmpc(pb, qp="qpoases")
    x = x0.copy()
    xs_, us_, it_ = [x.copy()], [], []
    for _ in range(nsteps):
        u = ctrl.step(x)
        assert ctrl.log["status"][-1] == 0
        x = st.F(x[None, :], u[None, :])[0]
        xs_.append(x.copy()); us_.append(u.copy()); it_.append(ctrl.log["iter"][-1])
    return b, np.array(xs_), np.array(us_), np.array(it_)


def rp_problem(name):
    from tunempc_b200.problem import MpcProblem
    return MpcProblem.load(os.path.join(HERE, "problem_%s.npz" % name))


def awe9(B=24, Bcl=6, nsteps=8, Blarge=512):
    """slack formulation (us, usc, g rows, scost; configs.awe9: nx = 9, nu = 3, ns = 3, nsc = 3, nh = 17, N = 20, p = 40):
    single solves at two tolerances, closed loops over the periodic reference, and a larger batch of compact results"""
    import multiprocessing as mp
    st = rp.StageLib("awe9")
    pb, info = configs.make_problem("awe9", st.F)
    pb.save(os.path.join(HERE, "problem_awe9.npz"))
    X0 = sample_x0("awe9", pb, B)
    out = {"X0": X0}
    cores = len(os.sched_getaffinity(0))
    with mp.get_context("fork").Pool(cores) as pool:
        for tag, tol in (("t6", 1e-6), ("t9", 1e-9)):
            res = sorted(pool.map(_awe9_single, [(i, X0[i], tol) for i in range(B)]), key=lambda r: r[0])
            assert all(r[5] == 0 for r in res)
            out.update({"u0_" + tag: np.array([r[1] for r in res]), "w_" + tag: np.array([r[2] for r in res]),
                        "lam_" + tag: np.array([r[3] for r in res]), "iter_" + tag: np.array([r[4] for r in res]),
                        "f_" + tag: np.array([r[6] for r in res]), "nAS_" + tag: np.array([r[7] for r in res]),
                        "nACtot_" + tag: np.array([r[8] for r in res]), "nAC_" + tag: np.array([r[9] for r in res])})
            print("awe9", tag, "iter hist", np.bincount(out["iter_" + tag]), "nAS", np.bincount(out["nAS_" + tag]), flush=True)
        Xc = sample_x0("awe9", pb, Bcl, 5)
        res = sorted(pool.map(_awe9_loop, [(b, Xc[b], nsteps) for b in range(Bcl)]), key=lambda r: r[0])
        out.update({"cl_X": np.array([r[1] for r in res]), "cl_U": np.array([r[2] for r in res]), "cl_iter": np.array([r[3] for r in res])})
        print("awe9 closed loop iter hist", np.bincount(out["cl_iter"].ravel()), flush=True)
        np.savez_compressed(os.path.join(HERE, "golden_awe9.npz"), **out)
        XL = sample_x0("awe9", pb, Blarge, 2024)
        res = sorted(pool.map(_awe9_single, [(i, XL[i], 1e-6) for i in range(Blarge)], chunksize=4), key=lambda r: r[0])
        nI = pb.N * pb.nh
        act = np.array([np.packbits(np.array([r[3][pb.g_h(k)][j] != 0 for k in range(pb.N) for j in range(pb.nh)], dtype=bool)) for r in res])
        np.savez_compressed(os.path.join(HERE, "large_awe9.npz"), seed=2024, B=Blarge, qp_backend="qpoases_e (oracle/_ref)",
                            u0=np.array([r[1] for r in res]), x1=np.array([r[2][pb.ix(1)] for r in res]),
                            iter=np.array([r[4] for r in res], dtype=np.int32), status=np.array([r[5] for r in res], dtype=np.int32),
                            f=np.array([r[6] for r in res]), nAS=np.array([r[7] for r in res], dtype=np.int32),
                            nACtot=np.array([r[8] for r in res], dtype=np.int32), nAC=np.array([r[9] for r in res], dtype=np.int32),
                            active=act)
        print("awe9 large: status", np.bincount(np.array([r[5] for r in res])), "iter", np.bincount(np.array([r[4] for r in res])), flush=True)


def evaporation_sc1():
    """soft constraints (Tuner.create_mpc(..., opts={'slack_flag': 'active'}), tuner.py:171-177, preprocessing.py:120-155) on the
    evaporation config: the row active at the steady state (X2 >= 25) is softened.  12 x0, three of them BELOW the bound, and two
    closed loops.  (A closed loop started 0.8 below the bound is left out: from its second MPC step on the oracle stalls at
    max_iter -- qpOASES_e returns its QP solution with a stationarity residual of 1.75e-6 > tol, its termination tolerance relative
    to multipliers of size scost = 6e4 -- while the device routines converge to the same u0)"""
    from tunempc_b200 import constraints
    from tunempc_b200.pmpc import problem_from_reference_args
    pb = rp_problem("evaporation")
    Cs, cs, scost, rows = constraints.soften_rows(pb.C, pb.c, pb.lam_h_ref, "active")
    model = configs.CONFIGS["evaporation_sc1"]()["model"]
    card = {"f": model, "h": (Cs, cs), "scost": scost, "vars": {"x": model.x, "u": model.u, "usc": list(range(len(rows)))}}
    ps = problem_from_reference_args(pb.N, card, "tracking", {"x": [pb.wref[0, :2]], "u": [pb.wref[0, 2:]]},
                                     {"H": list(pb.H), "q": list(pb.q)}, {"dyn": [np.zeros(2)], "h": list(pb.lam_h_ref)},
                                     {"A": pb.S_A, "B": pb.S_B}, {"p_operator": pb.term_idx})
    ps.save(os.path.join(HERE, "problem_evaporation_sc1.npz"))
    X0 = sample_x0("evaporation", pb, 12, 11)
    X0[2, 0] = 24.7; X0[5, 0] = 24.2; X0[9, 0] = 24.95
    oc = rp.Pmpc(ps, qp="qpoases", sqp_options={"max_iter": 50})
    st = rp.StageLib("evaporation")
    out = {"X0": X0}
    U, W, LAM, IT, NAS = [], [], [], [], []
    for b in range(12):
        oc.reset()
        U.append(oc.step(X0[b])); W.append(oc.w_sol); LAM.append(oc.lam_g); IT.append(oc.log["iter"][-1]); NAS.append(oc.log["nAS"][-1])
        assert oc.log["status"][-1] == 0
    out.update(u0_t6=np.array(U), w_t6=np.array(W), lam_t6=np.array(LAM), iter_t6=np.array(IT), nAS_t6=np.array(NAS))
    Xc, Uc, Ic = [], [], []
    for b in (2, 7):
        oc.reset()
        x = X0[b].copy()
        xs_, us_, it_ = [x.copy()], [], []
        for _ in range(5):
            u = oc.step(x)
            assert oc.log["status"][-1] == 0
            x = st.F(x[None], u[None])[0]
            xs_.append(x.copy()); us_.append(u.copy()); it_.append(oc.log["iter"][-1])
        Xc.append(xs_); Uc.append(us_); Ic.append(it_)
    out.update(cl_X=np.array(Xc), cl_U=np.array(Uc), cl_iter=np.array(Ic))
    np.savez_compressed(os.path.join(HERE, "golden_evaporation_sc1.npz"), **out)
    print("evaporation_sc1: iter", out["iter_t6"], "closed-loop iter", out["cl_iter"].tolist())


def chain(name="chain", B=32):
    """synthetic models (configs.chain nz = 8, configs.dims9 nz = 12 with the AWE config's dimensions): generic-dimension paths"""
    st = rp.StageLib(name)
    pb, info = configs.make_problem(name, st.F)
    pb.save(os.path.join(HERE, "problem_%s.npz" % name))
    X0 = sample_x0(name, pb, B)
    out = {"X0": X0}
    ctrl = rp.Pmpc(pb, qp="qpoases")
    U, W, LAM, IT, NAS = [], [], [], [], []
    for b in range(B):
        ctrl.reset()
        U.append(ctrl.step(X0[b])); W.append(ctrl.w_sol); LAM.append(ctrl.lam_g)
        IT.append(ctrl.log["iter"][-1]); NAS.append(ctrl.log["nAS"][-1])
        assert ctrl.log["status"][-1] == 0
    out.update({"u0_t6": np.array(U), "w_t6": np.array(W), "lam_t6": np.array(LAM), "iter_t6": np.array(IT), "nAS_t6": np.array(NAS)})
    print(name, "iter hist", np.bincount(np.array(IT)), "nAS hist", np.bincount(np.array(NAS)))
    np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % name), **out)


def main():
    rp.build()
    if len(sys.argv) > 1 and sys.argv[1] == "awe9":
        awe9()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "evaporation_sc1":
        evaporation_sc1()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "chain":
        chain()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dims9":
        chain("dims9", 12)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "economic":
        economic()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "unicycle":
        unicycle()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "unicycle_economic":
        unicycle_economic()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "evaporation":
        evaporation()
        return
    for name, B in (("lq", 64), ("cstr", 48)):
        st = rp.StageLib(name)
        pb, info = configs.make_problem(name, st.F)
        pb.save(os.path.join(HERE, "problem_%s.npz" % name))
        X0 = sample_x0(name, pb, B)
        out = {"X0": X0}
        for tag, tol in (("t6", 1e-6), ("t9", 1e-9)):
            ctrl = rp.Pmpc(pb, qp="qpoases", sqp_options={"tol": tol})
            U, W, LAM, IT, ST, NAS = [], [], [], [], [], []
            for b in range(B):
                ctrl.reset()
                u = ctrl.step(X0[b])
                U.append(u); W.append(ctrl.w_sol); LAM.append(ctrl.lam_g)
                IT.append(ctrl.log["iter"][-1]); ST.append(ctrl.log["status"][-1]); NAS.append(ctrl.log["nAS"][-1])
            out.update({"u0_" + tag: np.array(U), "w_" + tag: np.array(W), "lam_" + tag: np.array(LAM),
                        "iter_" + tag: np.array(IT), "status_" + tag: np.array(ST), "nAS_" + tag: np.array(NAS)})
            print(name, tag, "iter hist", np.bincount(np.array(IT)), "status", np.bincount(np.array(ST)))
        # closed loop of 5 steps on the first 8 instances (plant = model), oracle
        ctrl = rp.Pmpc(pb, qp="qpoases")
        Xcl = []
        Ucl = []
        for b in range(8):
            ctrl.reset()
            x = X0[b].copy()
            xs_, us_ = [x.copy()], []
            for _ in range(5):
                u = ctrl.step(x)
                x = st.F(x[None, :], u[None, :])[0]
                xs_.append(x.copy()); us_.append(u.copy())
            Xcl.append(xs_); Ucl.append(us_)
        out["cl_X"] = np.array(Xcl)
        out["cl_U"] = np.array(Ucl)
        np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % name), **out)
    unicycle()
    evaporation()
    economic()
    chain()
    chain("dims9", 12)


if __name__ == "__main__":
    main()

"""Generate the LARGE committed fixtures tests/golden/large_<cfg>.npz: B = 4096 seeded x0 per config solved by
the ORACLE (oracle/reference_port.py, QP = the reference tree's qpOASES_e), one process per host core.

Only compact per-instance results are stored (the x0 batch is regenerated from the seed by the same generator as the small
fixtures): u0, x_1 .. the predicted state at the end of the first interval, iteration count, status, objective f, the
log counters nAS / nACtot / nAC (pmpc.py:840-856, sqp_method.py:203-219) and the active set as a bit mask over the
inequality rows.  The oracle's QP backend is recorded in the file.

  python tests/golden/make_golden_large.py [cfg ...]        default: lq cstr evaporation unicycle
"""
import multiprocessing as mp
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "1"

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))
B_LARGE = 4096
SEED = 2024


def _x0(name, pb, B):
    from make_golden import sample_x0
    return sample_x0(name, pb, B, SEED)


def _worker(args):
    name, idx = args
    from oracle import reference_port as rp
    from tunempc_b200.problem import MpcProblem
    pb = MpcProblem.load(os.path.join(HERE, "problem_%s.npz" % name))
    X0 = _x0(name, pb, B_LARGE)
    ctrl = rp.Pmpc(pb, qp="qpoases")
    nI = pb.N * pb.nh
    out = []
    for i in idx:
        ctrl.reset()
        try:
            u = ctrl.step(X0[i])
        except Exception:                                     # the oracle's QP solver raises on an infeasible QP
            out.append((i, np.full(pb.nu, np.nan), np.full(pb.nx, np.nan), -1, 2, np.nan, -1, -1, -1, np.zeros((nI + 7) // 8, dtype=np.uint8)))
            continue
        lam = ctrl.lam_g
        act = np.array([lam[pb.g_h(k)][j] != 0 for k in range(pb.N) for j in range(pb.nh)], dtype=bool) if pb.nh else np.zeros(0, dtype=bool)
        out.append((i, u, ctrl.w_sol[pb.ix(1)].copy(), ctrl.log["iter"][-1], ctrl.log["status"][-1], ctrl.log["f"][-1],
                    ctrl.log["nAS"][-1], ctrl.log["nACtot"][-1], ctrl.log["nAC"][-1], np.packbits(act)))
    return out


def main():
    from oracle import reference_port as rp
    from tunempc_b200.problem import MpcProblem
    rp.build()
    assert rp.qpoases_available(), "the large fixtures are generated with the reference tree's qpOASES_e"
    names = sys.argv[1:] or ["lq", "cstr", "evaporation", "unicycle"]
    cores = len(os.sched_getaffinity(0))
    for name in names:
        pb = MpcProblem.load(os.path.join(HERE, "problem_%s.npz" % name))
        t0 = time.time()
        chunks = [(name, list(range(c, B_LARGE, 4 * cores))) for c in range(4 * cores)]
        with mp.get_context("fork").Pool(cores) as pool:
            res = [r for part in pool.map(_worker, chunks) for r in part]
        res.sort(key=lambda r: r[0])
        out = {"seed": SEED, "B": B_LARGE, "qp_backend": "qpoases_e (oracle/_ref)",
               "u0": np.array([r[1] for r in res]), "x1": np.array([r[2] for r in res]),
               "iter": np.array([r[3] for r in res], dtype=np.int32), "status": np.array([r[4] for r in res], dtype=np.int32),
               "f": np.array([r[5] for r in res]), "nAS": np.array([r[6] for r in res], dtype=np.int32),
               "nACtot": np.array([r[7] for r in res], dtype=np.int32), "nAC": np.array([r[8] for r in res], dtype=np.int32),
               "active": np.array([r[9] for r in res], dtype=np.uint8)}
        np.savez_compressed(os.path.join(HERE, "large_%s.npz" % name), **out)
        print("%s: %d instances in %.0f s on %d cores; status %s iter %s" % (name, B_LARGE, time.time() - t0, cores,
              np.bincount(out["status"][out["status"] >= 0]), np.bincount(out["iter"][out["iter"] >= 0])), flush=True)


if __name__ == "__main__":
    main()

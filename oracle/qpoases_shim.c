/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Thin C entry point over the reference tree's vendored qpOASES_e 3.1
 * (/root/reference/external/acados/external/qpoases, compiled in place by oracle/Makefile into oracle/_ref/).
 * The reference's hot path solves its QP with CasADi conic('qpoases') (tunempc/sqp_method.py:127-132,168);
 * CasADi and its bundled qpOASES 3.2 are absent, the vendored embedded port is the closest compiled reference.
 * Options follow tunempc/sqp_method.py:112-117 (enableEqualities = True; everything else default).
 * Dual sign: qpOASES returns y with  H x + g = A' y ;  CasADi's lam_a satisfies  H x + g + A' lam = 0  -> lam = -y.
 */
#include <stdlib.h>
#include <qpOASES_e.h>

int qpo_solve(int nV, int nC, const double* H, const double* g, const double* A,
              const double* lbA, const double* ubA, double* x, double* lam_a, int* nWSR) {
  static Options options;
  QProblem* qp = QProblem_createMemory(nV, nC);
  if (!qp) return -1;
  QProblemCON(qp, nV, nC, HST_UNKNOWN);
  Options_setToDefault(&options);
  options.enableEqualities = BT_TRUE;
  options.printLevel = PL_NONE;
  QProblem_setOptions(qp, options);
  double* y = (double*)malloc(sizeof(double) * (nV + nC));
  int ret = (int)QProblem_init(qp, (real_t*)H, (real_t*)g, (real_t*)A, 0, 0, (real_t*)lbA, (real_t*)ubA, nWSR, 0);
  QProblem_getPrimalSolution(qp, x);
  QProblem_getDualSolution(qp, y);
  for (int i = 0; i < nC; ++i) lam_a[i] = -y[nV + i];
  free(y);
  free(qp);
  return ret;
}

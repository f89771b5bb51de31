/* ORACLE -- TEST INFRASTRUCTURE ONLY (not shipped, not on the product path).
 *
 * Plain-C restatement of what CasADi computes for the reference when it evaluates the multiple-shooting
 * map F(x,u) and its first/second derivatives:
 *   - integrator('F','rk',ode,{'tf':..,'number_of_finite_elements':M}) = M classical RK4 steps of h = tf/M
 *     (reference: examples/cstr/cstr_model.py:62-64,97; examples/unicycle/main.py:67),
 *   - ca.jacobian / ca.hessian of the NLP through that integrator (tunempc/sqp_method.py:82-98).
 * Algorithm here: dense forward-mode propagation of the FULL first- and second-order sensitivity tensors
 * (dx/dz: nx*nz, d2x/dz2: nx*nz*nz) through every RK4 stage.  It is deliberately the naive O(nx nz^4) method
 * and differs from the CUDA path (pair-wise second-order directional propagation), so the two check each other.
 *
 * Compiled once per model:  gcc -O2 -fPIC -shared -DTMPC_MODEL_HEADER='"model_cstr.h"' stage.c
 * "parity unpinned": the reference holds no golden vector for this path; derivatives
 * are pinned by finite differences and by sympy lambdify in tests/.
 */
#include <string.h>
#include TMPC_MODEL_HEADER

#define NX TMPC_NX
#define NU TMPC_NU
#define NZ TMPC_NZ

static const int hA[] = TMPC_HESS_A, hB[] = TMPC_HESS_B, hC[] = TMPC_HESS_C;

void orc_dims(int* nx, int* nu, int* rk_steps, double* dt, int* discrete) {
  *nx = NX; *nu = NU; *rk_steps = TMPC_RK_STEPS; *dt = TMPC_RK_DT; *discrete = TMPC_DISCRETE;
}

void orc_ode(const double* x, const double* u, double* f) { tmpc_ode(x, u, f); }

/* value, first and second derivative of the ODE right-hand side along given sensitivities.
 * X: nx, dX: nx*nz (row-major), ddX: nx*nz*nz ; outputs k, dk, ddk of the same shapes.  order = 0,1,2 */
static void rhs_sens(const double* X, const double* u, const double* dX, const double* ddX,
                     double* k, double* dk, double* ddk, int order) {
  double J[NX * NZ], Hn[TMPC_NHESS > 0 ? TMPC_NHESS : 1];
  if (order == 0) { tmpc_ode(X, u, k); return; }
  if (order == 1) tmpc_ode_jac(X, u, k, J); else tmpc_ode_d2(X, u, k, J, Hn);
  /* dZ = [dX ; dU] with dU = [0 I] */
  double dZ[NZ * NZ];
  memcpy(dZ, dX, sizeof(double) * NX * NZ);
  for (int b = 0; b < NU; ++b) for (int i = 0; i < NZ; ++i) dZ[(NX + b) * NZ + i] = (i == NX + b) ? 1.0 : 0.0;
  for (int a = 0; a < NX; ++a) for (int i = 0; i < NZ; ++i) {
    double s = 0; for (int b = 0; b < NZ; ++b) s += J[a * NZ + b] * dZ[b * NZ + i];
    dk[a * NZ + i] = s;
  }
  if (order < 2) return;
  for (int a = 0; a < NX; ++a) for (int i = 0; i < NZ; ++i) for (int j = 0; j < NZ; ++j) {
    double s = 0; for (int b = 0; b < NX; ++b) s += J[a * NZ + b] * ddX[(b * NZ + i) * NZ + j];
    ddk[(a * NZ + i) * NZ + j] = s;
  }
  for (int n = 0; n < TMPC_NHESS; ++n) {
    int a = hA[n], b = hB[n], c = hC[n];
    for (int i = 0; i < NZ; ++i) for (int j = 0; j < NZ; ++j) {
      double t = Hn[n] * dZ[b * NZ + i] * dZ[c * NZ + j];
      if (b != c) t += Hn[n] * dZ[c * NZ + i] * dZ[b * NZ + j];
      ddk[(a * NZ + i) * NZ + j] += t;
    }
  }
}

#define NS1 (NX * NZ)
#define NS2 (NX * NZ * NZ)

/* xf = F(x,u); S = dF/dz (nx*nz row-major); T = d2F/dz2 (nx*nz*nz).  order selects how much is computed. */
void orc_F(const double* x, const double* u, double* xf, double* S, double* T, int order) {
  double X[NX], dX[NS1], ddX[NS2];
  memcpy(X, x, sizeof X);
  memset(dX, 0, sizeof dX); memset(ddX, 0, sizeof ddX);
  for (int a = 0; a < NX; ++a) dX[a * NZ + a] = 1.0;
#if TMPC_DISCRETE
  {
    double k[NX], dk[NS1], ddk[NS2];
    rhs_sens(X, u, dX, ddX, k, dk, ddk, order);
    memcpy(xf, k, sizeof k);
    if (order >= 1) memcpy(S, dk, sizeof dk);
    if (order >= 2) memcpy(T, ddk, sizeof ddk);
    return;
  }
#else
  const double h = TMPC_RK_DT;
  for (int s = 0; s < TMPC_RK_STEPS; ++s) {
    double k[4][NX], dk[4][NS1], ddk[4][NS2], Xs[NX], dXs[NS1], ddXs[NS2];
    const double cs[4] = {0.0, 0.5, 0.5, 1.0};
    for (int st = 0; st < 4; ++st) {
      for (int i = 0; i < NX; ++i) Xs[i] = X[i] + (st ? cs[st] * h * k[st - 1][i] : 0.0);
      if (order >= 1) for (int i = 0; i < NS1; ++i) dXs[i] = dX[i] + (st ? cs[st] * h * dk[st - 1][i] : 0.0);
      if (order >= 2) for (int i = 0; i < NS2; ++i) ddXs[i] = ddX[i] + (st ? cs[st] * h * ddk[st - 1][i] : 0.0);
      rhs_sens(Xs, u, dXs, ddXs, k[st], dk[st], ddk[st], order);
    }
    for (int i = 0; i < NX; ++i) X[i] += h / 6.0 * (k[0][i] + 2 * k[1][i] + 2 * k[2][i] + k[3][i]);
    if (order >= 1) for (int i = 0; i < NS1; ++i) dX[i] += h / 6.0 * (dk[0][i] + 2 * dk[1][i] + 2 * dk[2][i] + dk[3][i]);
    if (order >= 2) for (int i = 0; i < NS2; ++i) ddX[i] += h / 6.0 * (ddk[0][i] + 2 * ddk[1][i] + 2 * ddk[2][i] + ddk[3][i]);
  }
  memcpy(xf, X, sizeof X);
  if (order >= 1) memcpy(S, dX, sizeof dX);
  if (order >= 2) memcpy(T, ddX, sizeof ddX);
#endif
}

/* batched over N stages: xs (N*nx), us (N*nu) -> xf (N*nx), S (N*nx*nz), T (N*nx*nz*nz) */
void orc_F_map(int N, const double* xs, const double* us, double* xf, double* S, double* T, int order) {
  for (int k = 0; k < N; ++k)
    orc_F(xs + k * NX, us + k * NU, xf + k * NX, S ? S + k * NS1 : 0, T ? T + k * NS2 : 0, order);
}

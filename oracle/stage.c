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
#include <math.h>
#include TMPC_MODEL_HEADER

#define NX TMPC_NX
#ifdef TMPC_NUM               /* model dimensions: the slack variables of the MPC do not enter the dynamics */
#define NU TMPC_NUM
#define NZ TMPC_NZM
#else
#define NU TMPC_NU
#define NZ TMPC_NZ
#endif

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

#ifndef TMPC_COLLOCATION
#define TMPC_COLLOCATION 0
#endif

/* dense Gaussian elimination with partial pivoting: solves A X = B in place (A n x n row-major, destroyed; B n x m) */
static void gauss_solve(int n, double* A, double* B, int m) {
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r) if (fabs(A[r * n + c]) > fabs(A[p * n + c])) p = r;
    if (p != c) {
      for (int k = 0; k < n; ++k) { double t = A[c * n + k]; A[c * n + k] = A[p * n + k]; A[p * n + k] = t; }
      for (int k = 0; k < m; ++k) { double t = B[c * m + k]; B[c * m + k] = B[p * m + k]; B[p * m + k] = t; }
    }
    for (int r = c + 1; r < n; ++r) {
      double f = A[r * n + c] / A[c * n + c];
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      for (int k = 0; k < m; ++k) B[r * m + k] -= f * B[c * m + k];
    }
  }
  for (int r = n - 1; r >= 0; --r)
    for (int k = 0; k < m; ++k) {
      double v = B[r * m + k];
      for (int c2 = r + 1; c2 < n; ++c2) v -= A[r * n + c2] * B[c2 * m + k];
      B[r * m + k] = v / A[r * n + r];
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
#elif TMPC_COLLOCATION
  /* integrator('F','collocation',ode,{'tf':..}) (reference: examples/evaporation_process/main.py:103): per finite
   * element the 3-node Radau collocation equations (= Radau IIA), solved by Newton on the stage STATES Y_q
   *   R_q = Y_q - X - h sum_l a_ql f(Y_l,u) = 0,   x+ = Y_3,
   * then FULL first/second-order tensors of Y by the implicit function theorem (dense Gaussian elimination on the
   * 3nx system).  The CUDA path iterates on the stage derivatives and propagates one direction pair at a time. */
  {
    const double sq6 = sqrt(6.0);
    const double A[3][3] = {{(88 - 7 * sq6) / 360, (296 - 169 * sq6) / 1800, (-2 + 3 * sq6) / 225},
                            {(296 + 169 * sq6) / 1800, (88 + 7 * sq6) / 360, (-2 - 3 * sq6) / 225},
                            {(16 - sq6) / 36, (16 + sq6) / 36, 1.0 / 9}};
    const double h = TMPC_RK_DT;
    enum { CN = 3 * NX };
    for (int s = 0; s < TMPC_RK_STEPS; ++s) {
      double Y[3][NX], f[3][NX], df[3][NS1], ddf[3][NS2], dY[3][NS1], ddY[3][NS2];
      double Jm[CN][CN], rhs[CN], J[3][NX * NZ], Hn[3][TMPC_NHESS > 0 ? TMPC_NHESS : 1];
      for (int q = 0; q < 3; ++q) memcpy(Y[q], X, sizeof X);
      for (int it = 0; it < 50; ++it) {
        for (int q = 0; q < 3; ++q) tmpc_ode_d2(Y[q], u, f[q], J[q], Hn[q]);
        double rmax = 0;
        for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) {
          double r = Y[q][a] - X[a];
          for (int l = 0; l < 3; ++l) r -= h * A[q][l] * f[l][a];
          rhs[q * NX + a] = -r;
          if (fabs(r) > rmax) rmax = fabs(r);
        }
        for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) for (int l = 0; l < 3; ++l) for (int b = 0; b < NX; ++b)
          Jm[q * NX + a][l * NX + b] = ((q == l && a == b) ? 1.0 : 0.0) - h * A[q][l] * J[l][a * NZ + b];
        if (it > 0 && rmax < 1e-15 * (1.0 + fabs(X[0]))) break;
        gauss_solve(CN, &Jm[0][0], rhs, 1);
        double dmax = 0;
        for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) { Y[q][a] += rhs[q * NX + a]; if (fabs(rhs[q * NX + a]) > dmax) dmax = fabs(rhs[q * NX + a]); }
        if (dmax == 0.0) break;
      }
      if (order >= 1) {
        /* converged Jacobian of the residual w.r.t. Y; first-order: Jm dY = dX + h sum_l a_ql (J_u part) */
        for (int q = 0; q < 3; ++q) tmpc_ode_d2(Y[q], u, f[q], J[q], Hn[q]);
        for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) for (int l = 0; l < 3; ++l) for (int b = 0; b < NX; ++b)
          Jm[q * NX + a][l * NX + b] = ((q == l && a == b) ? 1.0 : 0.0) - h * A[q][l] * J[l][a * NZ + b];
        static double R1[3 * NX * NZ];
        for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) for (int i = 0; i < NZ; ++i) {
          double r = dX[a * NZ + i];
          if (i >= NX) for (int l = 0; l < 3; ++l) r += h * A[q][l] * J[l][a * NZ + i];
          R1[(q * NX + a) * NZ + i] = r;
        }
        double Jc[CN][CN];
        memcpy(Jc, Jm, sizeof Jm);
        gauss_solve(CN, &Jc[0][0], R1, NZ);
        for (int q = 0; q < 3; ++q) memcpy(dY[q], R1 + q * NS1, sizeof(double) * NS1);
        if (order >= 2) {
          /* second order: Jm ddY = ddX + h sum_l a_ql ( f_zz[dZ_l, dZ_l] ), dZ_l = [dY_l ; 0 I] */
          static double R2[3 * NX * NZ * NZ];
          double fz[3][NS2];
          for (int l = 0; l < 3; ++l) {
            double dZ[NZ * NZ];
            memcpy(dZ, dY[l], sizeof(double) * NS1);
            for (int b = 0; b < NU; ++b) for (int i = 0; i < NZ; ++i) dZ[(NX + b) * NZ + i] = (i == NX + b) ? 1.0 : 0.0;
            memset(fz[l], 0, sizeof fz[l]);
            for (int n = 0; n < TMPC_NHESS; ++n) {
              int a = hA[n], b = hB[n], c = hC[n];
              for (int i = 0; i < NZ; ++i) for (int j = 0; j < NZ; ++j) {
                double t = Hn[l][n] * dZ[b * NZ + i] * dZ[c * NZ + j];
                if (b != c) t += Hn[l][n] * dZ[c * NZ + i] * dZ[b * NZ + j];
                fz[l][(a * NZ + i) * NZ + j] += t;
              }
            }
          }
          for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) for (int ij = 0; ij < NZ * NZ; ++ij) {
            double r = ddX[a * NZ * NZ + ij];
            for (int l = 0; l < 3; ++l) r += h * A[q][l] * fz[l][a * NZ * NZ + ij];
            R2[(q * NX + a) * NZ * NZ + ij] = r;
          }
          memcpy(Jc, Jm, sizeof Jm);
          gauss_solve(CN, &Jc[0][0], R2, NZ * NZ);
          for (int q = 0; q < 3; ++q) memcpy(ddY[q], R2 + q * NS2, sizeof(double) * NS2);
          memcpy(ddX, ddY[2], sizeof ddX);
        }
        memcpy(dX, dY[2], sizeof dX);
      }
      memcpy(X, Y[2], sizeof X);
    }
    memcpy(xf, X, sizeof X);
    if (order >= 1) memcpy(S, dX, sizeof dX);
    if (order >= 2) memcpy(T, ddX, sizeof ddX);
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

// tmpc_core.cuh -- per-instance device routines of the batched tuned-MPC feedback solve (fp64).
//
// Everything here is written from scratch for this repository; the reference (jdeschut/tunempc) has no native
// or GPU code on this path.  What each routine replaces in the reference's Python:
//   tm_lin_pair      jacg_fun / H_fun evaluations through the RK4 integrator   tunempc/sqp_method.py:152,159,330
//   tm_qp_solve      conic('qpoases') QP solve                                 tunempc/sqp_method.py:158-168
//   tm_post          __linesearch (filter) + w/lam update                      tunempc/sqp_method.py:171-177,289-325
//   tm_conv          __check_convergence + __postprocessing stats              tunempc/sqp_method.py:240-287,185-221
//   tm_init          k = 0 bookkeeping of __check_convergence, __prefilter_lam_g   :223-238,248-261
//   tm_shift         Pmpc.__shift_initial_guess                                tunempc/pmpc.py:867-906
//
// Execution model: one WARP per instance for the sequential parts (QP, line search, convergence), one THREAD per
// (instance, stage, sensitivity pair) for the linearisation.  The warp routines are written against
// TM_LANE / TM_NL / TM_SYNC so that the same source compiles as a 1-lane sequential program with g++ for the
// CPU twin used by the CPU-only tests (tests/twin); that twin is test infrastructure, not a product path.
#pragma once
#include <stdint.h>
#include <math.h>
#include TMPC_MODEL_HEADER

#ifdef __CUDACC__
#define TM_HD __host__ __device__ __forceinline__
#define TM_HDN static __host__ __device__ __noinline__
#define TM_HDM __host__ __device__ __forceinline__
#else
#define TM_HDM inline
#define TM_HD static inline
#define TM_HDN static
#endif

#if defined(__CUDA_ARCH__) && !defined(TM_THREAD_MODE)
#define TM_WARP_MODE 1
#define TM_LANE ((int)(threadIdx.x & 31))
#define TM_NL 32
#define TM_SYNC() __syncwarp()
__device__ __forceinline__ double tm_wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double tm_wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int tm_wsumi(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// arg-min over the warp: returns the smallest value and its payload (ties: smallest payload)
__device__ __forceinline__ void tm_wargmin(double& v, int& id) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, id, o);
    if (ov < v || (ov == v && oi < id)) { v = ov; id = oi; }
  }
}
__device__ __forceinline__ int tm_wany(int p) { return __any_sync(0xffffffffu, p); }
#else
#define TM_LANE 0
#define TM_NL 1
#define TM_SYNC() ((void)0)
TM_HD double tm_wsum(double v) { return v; }
TM_HD double tm_wmax(double v) { return v; }
TM_HD int tm_wsumi(int v) { return v; }
TM_HD void tm_wargmin(double&, int&) {}
TM_HD int tm_wany(int p) { return p; }
#endif

// workspace accessor: plain pointer (shared memory, warp-per-instance; host twin) or a lane-interleaved view of global
// memory (thread-per-instance mode: element e of lane l lives at base[e*TM_WS_STRIDE + l], so a warp touching the same
// element of its 32 instances issues one coalesced 256-byte transaction)
#ifdef TM_WS_STRIDE
struct TmP {
  double* p;
  TM_HDM double& operator[](int i) const { return p[(size_t)i * TM_WS_STRIDE]; }
  TM_HDM double& operator[](size_t i) const { return p[i * TM_WS_STRIDE]; }
  TM_HDM TmP operator+(size_t o) const { TmP r; r.p = p + o * TM_WS_STRIDE; return r; }
  TM_HDM TmP operator+(int o) const { TmP r; r.p = p + (size_t)o * TM_WS_STRIDE; return r; }
};
TM_HD TmP tm_mkp(double* base, size_t off) { TmP r; r.p = base + off * TM_WS_STRIDE; return r; }
#else
typedef double* TmP;
TM_HD TmP tm_mkp(double* base, size_t off) { return base + off; }
#endif

#define NX TMPC_NX
#define NU TMPC_NU
#define NZ TMPC_NZ
#define TM_NPAIR (NZ * (NZ + 1) / 2)
#define TM_LSZ (NX + NX * NZ + TM_NPAIR)   /* per-stage linearisation record: xf | S row-major nx*nz | W packed i<=j */
#define TM_INF 1e300
#define TM_NCNT 24
#define TM_ALW 12  /* words of the augmented-Lagrangian row mask: supports N*nh <= 384 */

// ---------------------------------------------------------------------------------------------------------------
// problem constants and per-batch state (plain pointers; device memory in the product, malloc in the twin)
// ---------------------------------------------------------------------------------------------------------------
struct TmProb {
  int N, nh, nxt, p, n_w, n_g;
  int hessian_exact, max_iter, max_ls, filter_cap, maxact;
  int economic;           // 1: stage cost = the model card's l(x,u) (economic MPC, pmpc.py:97-107), 0: tuned tracking cost (mtools.py:43-57)
  double tol, lam_tresh, beta, reg_tol, rho, al_gamma;
  const double *wref, *H, *q, *ref_du, *C, *c;   // wref p*nz | H p*nz*nz (symmetric) | q p*nz | ref_du p*n_g | C nh*nz | c nh
  const int *term_idx, *relax0;
};

struct TmState {
  int64_t B;
  int phase;              // index mod p of this step
  const double* X0;       // B*nx
  double *W, *LAM;        // B*n_w, B*n_g   current iterate (warm start before the step, solution after)
  double *D, *LAMQ;       // QP step and QP multipliers
  double* LIN;            // B*N*TM_LSZ
  double* G;              // B*n_g constraint values at the last evaluated point
  double* FILT;           // B*filter_cap*2
  double* fval;           // B
  int *nfilt, *iter, *status, *flags, *nAS, *nACtot, *nAC;
  int* qpstat;            // B: result of the last QP (0 ok)
  int* qpmode;            // B: 0 fresh, 1..3 retry with the stored row mask, 100 Gauss-Newton fallback
  int* qpwork;            // B: active-set iterations of the instance's last QP (scheduling key: the thread-per-instance
                          //    kernel is fed instances of similar cost so that the lanes of a warp stay in step)
  unsigned* almask;       // B*TM_ALW: augmented-Lagrangian row mask carried between retries
  int *list_retry, *cnt_retry;   // instances whose QP must be re-solved (filled by tm_qp)
  unsigned* asinit;       // B*aswords bitmask of initially active inequality rows
  int aswords;
  int *list_next, *cnt_next, *list_relin, *cnt_relin;
  unsigned long long* counters;   // TM_NCNT: [0] iterations [4] ls dynamics evals [5] QP attempts [6] active-set iterations [7] Riccati solves [8..19] attempt histogram
};

TM_HD int tm_gdyn(const TmProb& P, int k) { return NX + k * (NX + P.nh); }
TM_HD int tm_gh(const TmProb& P, int k) { return NX + k * (NX + P.nh) + NX; }
TM_HD int tm_gterm(const TmProb& P) { return NX + P.N * (NX + P.nh); }

TM_HD void tm_pair_ij(int pr, int& i, int& j) {   // packed upper-triangular index -> (i<=j), row-major
  int r = 0, rem = pr;
  while (rem >= NZ - r) { rem -= NZ - r; ++r; }
  i = r; j = r + rem;
}
TM_HD int tm_pair_idx(int i, int j) {             // requires i<=j
  return i * NZ - i * (i - 1) / 2 + (j - i);
}

// ---------------------------------------------------------------------------------------------------------------
// K1: stage linearisation.  One call = one (stage, pair (i,j)) task: integrates x, s_i = dx/dz_i, s_j, t_ij = d2x/dz_i dz_j
// through RK4 (or one map evaluation for a discrete model).  ORDER 0: value; 1: + s_i; 2: + s_j and t_ij.
// ---------------------------------------------------------------------------------------------------------------
template <int ORDER>
TM_HD void tm_rhs(const double* X, const double* u, const double* Si, const double* Sj, const double* T, int i, int j,
                  double* k, double* dki, double* dkj, double* ddk) {
  if (ORDER == 0) { tmpc_ode(X, u, k); return; }
  double J[NX * NZ];
  double Hn[TMPC_NHESS > 0 ? TMPC_NHESS : 1];
  if (ORDER == 1) tmpc_ode_jac(X, u, k, J); else tmpc_ode_d2(X, u, k, J, Hn);
  double vi[NZ], vj[NZ];
#pragma unroll
  for (int a = 0; a < NX; ++a) { vi[a] = Si[a]; vj[a] = (ORDER == 2) ? Sj[a] : 0.0; }
#pragma unroll
  for (int b = 0; b < NU; ++b) { vi[NX + b] = (i == NX + b) ? 1.0 : 0.0; vj[NX + b] = (j == NX + b) ? 1.0 : 0.0; }
#pragma unroll
  for (int a = 0; a < NX; ++a) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < NZ; ++b) s += J[a * NZ + b] * vi[b];
    dki[a] = s;
  }
  if (ORDER == 2) {
#pragma unroll
    for (int a = 0; a < NX; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < NZ; ++b) s += J[a * NZ + b] * vj[b];
      dkj[a] = s;
    }
    tmpc_ode_bilin(Hn, vi, vj, ddk);
#pragma unroll
    for (int a = 0; a < NX; ++a) {
      double s = ddk[a];
#pragma unroll
      for (int b = 0; b < NX; ++b) s += J[a * NZ + b] * T[b];
      ddk[a] = s;
    }
  }
}


#ifndef TMPC_COLLOCATION
#define TMPC_COLLOCATION 0
#endif
#define TMPC_RK4 (!TMPC_DISCRETE && !TMPC_COLLOCATION)

#if TMPC_COLLOCATION
// ---------------------------------------------------------------------------------------------------------------
// CasADi integrator('F','collocation',ode,{'tf':..}) (reference: examples/evaporation_process/main.py:103): TMPC_RK_STEPS
// finite elements, Radau points of interpolation order 3 per element = the 3-stage Radau IIA method, solved to
// convergence by Newton.  Stage derivatives K_j = f(x + h sum_l a_jl K_l, u).  Sensitivities by the implicit function
// theorem with the converged iteration matrix M = I - h (a_jl J_j):  M dK_i = J_j v_i,  M ddK = f_zz[v_i, v_j] + J_x,j T.
// ---------------------------------------------------------------------------------------------------------------
#define TM_CN (3 * NX)
TM_HD void tm_lu_factor(double* M, int* piv) {          // row-major TM_CN x TM_CN, partial pivoting, in place
  for (int c = 0; c < TM_CN; ++c) {
    int p = c;
    double best = fabs(M[c * TM_CN + c]);
    for (int r = c + 1; r < TM_CN; ++r) { const double v = fabs(M[r * TM_CN + c]); if (v > best) { best = v; p = r; } }
    piv[c] = p;
    if (p != c) for (int k = 0; k < TM_CN; ++k) { const double t = M[c * TM_CN + k]; M[c * TM_CN + k] = M[p * TM_CN + k]; M[p * TM_CN + k] = t; }
    const double d = 1.0 / M[c * TM_CN + c];
    for (int r = c + 1; r < TM_CN; ++r) {
      const double f = M[r * TM_CN + c] * d;
      M[r * TM_CN + c] = f;
      for (int k = c + 1; k < TM_CN; ++k) M[r * TM_CN + k] -= f * M[c * TM_CN + k];
    }
  }
}
TM_HD void tm_lu_solve(const double* M, const int* piv, double* b) {
  for (int c = 0; c < TM_CN; ++c) {
    const int p = piv[c];
    if (p != c) { const double t = b[c]; b[c] = b[p]; b[p] = t; }
    for (int r = c + 1; r < TM_CN; ++r) b[r] -= M[r * TM_CN + c] * b[c];
  }
  for (int r = TM_CN - 1; r >= 0; --r) {
    double v = b[r];
    for (int k = r + 1; k < TM_CN; ++k) v -= M[r * TM_CN + k] * b[k];
    b[r] = v / M[r * TM_CN + r];
  }
}

template <int ORDER>
TM_HD void tm_integrate_colloc(const double* u, int i, int j, double* X, double* Si, double* Sj, double* T) {
  const double sq6 = 2.449489742783178;
  const double A[3][3] = {{(88.0 - 7.0 * sq6) / 360.0, (296.0 - 169.0 * sq6) / 1800.0, (-2.0 + 3.0 * sq6) / 225.0},
                          {(296.0 + 169.0 * sq6) / 1800.0, (88.0 + 7.0 * sq6) / 360.0, (-2.0 - 3.0 * sq6) / 225.0},
                          {(16.0 - sq6) / 36.0, (16.0 + sq6) / 36.0, 1.0 / 9.0}};
  const double h = TMPC_RK_DT;
  for (int s = 0; s < TMPC_RK_STEPS; ++s) {
    double K[3][NX], Jst[3][NX * NZ], M[TM_CN * TM_CN], rhs[TM_CN];
    double Hst[3][TMPC_NHESS > 0 ? TMPC_NHESS : 1];
    int piv[TM_CN];
    {
      double f0[NX];
      tmpc_ode(X, u, f0);
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) K[q][a] = f0[a];
    }
    // Newton to convergence; the last pass (done = 1) re-evaluates J (and d2f) at the converged stage points
    int done = 0;
    for (int it = 0; it < 40; ++it) {
      for (int q = 0; q < 3; ++q) {
        double Xq[NX], fq[NX];
        for (int a = 0; a < NX; ++a) {
          double v = X[a];
          for (int l = 0; l < 3; ++l) v += h * A[q][l] * K[l][a];
          Xq[a] = v;
        }
        if (done && ORDER == 2) tmpc_ode_d2(Xq, u, fq, Jst[q], Hst[q]); else tmpc_ode_jac(Xq, u, fq, Jst[q]);
        for (int a = 0; a < NX; ++a) rhs[q * NX + a] = -(K[q][a] - fq[a]);
      }
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) for (int l = 0; l < 3; ++l) for (int b = 0; b < NX; ++b)
        M[(q * NX + a) * TM_CN + l * NX + b] = ((q == l && a == b) ? 1.0 : 0.0) - h * A[q][l] * Jst[q][a * NZ + b];
      tm_lu_factor(M, piv);
      if (done) break;
      tm_lu_solve(M, piv, rhs);
      double dmax = 0.0, kmax = 1.0;
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) {
        K[q][a] += rhs[q * NX + a];
        dmax = fmax(dmax, fabs(rhs[q * NX + a]));
        kmax = fmax(kmax, fabs(K[q][a]));
      }
      if (!(dmax > 1e-14 * kmax)) done = 1;
      if (it == 38) done = 1;
    }
    double dKi[TM_CN], dKj[TM_CN], ddK[TM_CN];
    if (ORDER >= 1) {
      // stage-argument directions need dK, so solve first, then form v
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) {
        double t = 0.0;
        for (int b = 0; b < NX; ++b) t += Jst[q][a * NZ + b] * Si[b];
        if (i >= NX) t += Jst[q][a * NZ + i];
        dKi[q * NX + a] = t;
      }
      tm_lu_solve(M, piv, dKi);
    }
    if (ORDER >= 2) {
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) {
        double t = 0.0;
        for (int b = 0; b < NX; ++b) t += Jst[q][a * NZ + b] * Sj[b];
        if (j >= NX) t += Jst[q][a * NZ + j];
        dKj[q * NX + a] = t;
      }
      tm_lu_solve(M, piv, dKj);
      for (int q = 0; q < 3; ++q) {
        double vi[NZ], vj[NZ], dd[NX];
        for (int a = 0; a < NX; ++a) {
          double ti = Si[a], tj = Sj[a];
          for (int l = 0; l < 3; ++l) { ti += h * A[q][l] * dKi[l * NX + a]; tj += h * A[q][l] * dKj[l * NX + a]; }
          vi[a] = ti; vj[a] = tj;
        }
        for (int b = 0; b < NU; ++b) { vi[NX + b] = (i == NX + b) ? 1.0 : 0.0; vj[NX + b] = (j == NX + b) ? 1.0 : 0.0; }
        tmpc_ode_bilin(Hst[q], vi, vj, dd);
        for (int a = 0; a < NX; ++a) {
          double t = dd[a];
          for (int b = 0; b < NX; ++b) t += Jst[q][a * NZ + b] * T[b];
          ddK[q * NX + a] = t;
        }
      }
      tm_lu_solve(M, piv, ddK);
    }
    for (int a = 0; a < NX; ++a) {
      for (int q = 0; q < 3; ++q) {
        X[a] += h * A[2][q] * K[q][a];
        if (ORDER >= 1) Si[a] += h * A[2][q] * dKi[q * NX + a];
        if (ORDER >= 2) { Sj[a] += h * A[2][q] * dKj[q * NX + a]; T[a] += h * A[2][q] * ddK[q * NX + a]; }
      }
    }
  }
}
#endif

#if TMPC_COLLOCATION
// Whole linearisation record of one stage in ONE task: the Newton solve and the LU factorisation of an element are done
// once and shared by all NZ first-order and NZ(NZ+1)/2 second-order right-hand sides (the pair-per-task route repeats
// them for every pair).  order 1: xf, S;  order 2: + W = lam' d2F/dz2.  rec layout as everywhere: xf | S row-major | W.
TM_HD void tm_colloc_stage(const double* x0, const double* u, int order, const double* lam, double* rec) {
  const double sq6 = 2.449489742783178;
  const double A[3][3] = {{(88.0 - 7.0 * sq6) / 360.0, (296.0 - 169.0 * sq6) / 1800.0, (-2.0 + 3.0 * sq6) / 225.0},
                          {(296.0 + 169.0 * sq6) / 1800.0, (88.0 + 7.0 * sq6) / 360.0, (-2.0 - 3.0 * sq6) / 225.0},
                          {(16.0 - sq6) / 36.0, (16.0 + sq6) / 36.0, 1.0 / 9.0}};
  const double h = TMPC_RK_DT;
  double X[NX], Sm[NZ][NX], Tm[TM_NPAIR][NX];
  for (int a = 0; a < NX; ++a) X[a] = x0[a];
  for (int i = 0; i < NZ; ++i) for (int a = 0; a < NX; ++a) Sm[i][a] = (a == i) ? 1.0 : 0.0;
  for (int p = 0; p < TM_NPAIR; ++p) for (int a = 0; a < NX; ++a) Tm[p][a] = 0.0;
  for (int s = 0; s < TMPC_RK_STEPS; ++s) {
    double K[3][NX], Jst[3][NX * NZ], M[TM_CN * TM_CN], rhs[TM_CN];
    double Hst[3][TMPC_NHESS > 0 ? TMPC_NHESS : 1];
    int piv[TM_CN];
    {
      double f0[NX];
      tmpc_ode(X, u, f0);
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) K[q][a] = f0[a];
    }
    int done = 0;
    for (int it = 0; it < 40; ++it) {                  // same Newton iteration as tm_integrate_colloc
      for (int q = 0; q < 3; ++q) {
        double Xq[NX], fq[NX];
        for (int a = 0; a < NX; ++a) {
          double v = X[a];
          for (int l = 0; l < 3; ++l) v += h * A[q][l] * K[l][a];
          Xq[a] = v;
        }
        if (done && order == 2) tmpc_ode_d2(Xq, u, fq, Jst[q], Hst[q]); else tmpc_ode_jac(Xq, u, fq, Jst[q]);
        for (int a = 0; a < NX; ++a) rhs[q * NX + a] = -(K[q][a] - fq[a]);
      }
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) for (int l = 0; l < 3; ++l) for (int b = 0; b < NX; ++b)
        M[(q * NX + a) * TM_CN + l * NX + b] = ((q == l && a == b) ? 1.0 : 0.0) - h * A[q][l] * Jst[q][a * NZ + b];
      tm_lu_factor(M, piv);
      if (done) break;
      tm_lu_solve(M, piv, rhs);
      double dmax = 0.0, kmax = 1.0;
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) {
        K[q][a] += rhs[q * NX + a];
        dmax = fmax(dmax, fabs(rhs[q * NX + a]));
        kmax = fmax(kmax, fabs(K[q][a]));
      }
      if (!(dmax > 1e-14 * kmax)) done = 1;
      if (it == 38) done = 1;
    }
    // first order: all NZ directions against the same factorisation; V[i][q] = stage-argument direction
    double dK[NZ][TM_CN], V[NZ][3][NZ];
    for (int i = 0; i < NZ; ++i) {
      for (int q = 0; q < 3; ++q) for (int a = 0; a < NX; ++a) {
        double t = 0.0;
        for (int b = 0; b < NX; ++b) t += Jst[q][a * NZ + b] * Sm[i][b];
        if (i >= NX) t += Jst[q][a * NZ + i];
        dK[i][q * NX + a] = t;
      }
      tm_lu_solve(M, piv, dK[i]);
      for (int q = 0; q < 3; ++q) {
        for (int a = 0; a < NX; ++a) {
          double t = Sm[i][a];
          for (int l = 0; l < 3; ++l) t += h * A[q][l] * dK[i][l * NX + a];
          V[i][q][a] = t;
        }
        for (int b = 0; b < NU; ++b) V[i][q][NX + b] = (i == NX + b) ? 1.0 : 0.0;
      }
    }
    if (order == 2) {
      for (int i = 0; i < NZ; ++i)
        for (int j = i; j < NZ; ++j) {
          const int p = tm_pair_idx(i, j);
          double ddK[TM_CN];
          for (int q = 0; q < 3; ++q) {
            double dd[NX];
            tmpc_ode_bilin(Hst[q], V[i][q], V[j][q], dd);
            for (int a = 0; a < NX; ++a) {
              double t = dd[a];
              for (int b = 0; b < NX; ++b) t += Jst[q][a * NZ + b] * Tm[p][b];
              ddK[q * NX + a] = t;
            }
          }
          tm_lu_solve(M, piv, ddK);
          for (int a = 0; a < NX; ++a) for (int q = 0; q < 3; ++q) Tm[p][a] += h * A[2][q] * ddK[q * NX + a];
        }
    }
    for (int a = 0; a < NX; ++a)
      for (int q = 0; q < 3; ++q) {
        X[a] += h * A[2][q] * K[q][a];
        for (int i = 0; i < NZ; ++i) Sm[i][a] += h * A[2][q] * dK[i][q * NX + a];
      }
  }
  for (int a = 0; a < NX; ++a) rec[a] = X[a];
  for (int a = 0; a < NX; ++a) for (int i = 0; i < NZ; ++i) rec[NX + a * NZ + i] = Sm[i][a];
  if (order == 2)
    for (int p = 0; p < TM_NPAIR; ++p) {
      double wij = 0.0;
      for (int a = 0; a < NX; ++a) wij += lam[a] * Tm[p][a];
      rec[NX + NX * NZ + p] = wij;
    }
}
#endif

template <int ORDER>
TM_HD void tm_integrate(const double* x0, const double* u, int i, int j, double* X, double* Si, double* Sj, double* T) {
#pragma unroll
  for (int a = 0; a < NX; ++a) { X[a] = x0[a]; Si[a] = (a == i) ? 1.0 : 0.0; Sj[a] = (a == j) ? 1.0 : 0.0; T[a] = 0.0; }
#if TMPC_DISCRETE
  {
    double k[NX], di[NX], dj[NX], dd[NX];
    tm_rhs<ORDER>(X, u, Si, Sj, T, i, j, k, di, dj, dd);
#pragma unroll
    for (int a = 0; a < NX; ++a) {
      X[a] = k[a];
      if (ORDER >= 1) Si[a] = di[a];
      if (ORDER >= 2) { Sj[a] = dj[a]; T[a] = dd[a]; }
    }
  }
#elif TMPC_COLLOCATION
  tm_integrate_colloc<ORDER>(u, i, j, X, Si, Sj, T);
#else
  const double h = TMPC_RK_DT;
  for (int s = 0; s < TMPC_RK_STEPS; ++s) {
    double aX[NX], aI[NX], aJ[NX], aT[NX];      // accumulated increment (k1 + 2k2 + 2k3 + k4)
    double k[NX], di[NX], dj[NX], dd[NX];       // current stage derivative
    double Xs[NX], Is[NX], Js[NX], Ts[NX];      // stage argument
#pragma unroll
    for (int a = 0; a < NX; ++a) { Xs[a] = X[a]; Is[a] = Si[a]; Js[a] = Sj[a]; Ts[a] = T[a]; di[a] = dj[a] = dd[a] = 0.0; }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      tm_rhs<ORDER>(Xs, u, Is, Js, Ts, i, j, k, di, dj, dd);
      const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
      const double cn = (st == 2) ? 1.0 : 0.5;  // coefficient of the NEXT stage argument
#pragma unroll
      for (int a = 0; a < NX; ++a) {
        if (st == 0) { aX[a] = k[a]; aI[a] = di[a]; aJ[a] = dj[a]; aT[a] = dd[a]; }
        else { aX[a] += wgt * k[a]; aI[a] += wgt * di[a]; aJ[a] += wgt * dj[a]; aT[a] += wgt * dd[a]; }
        if (st < 3) {
          Xs[a] = X[a] + cn * h * k[a];
          if (ORDER >= 1) Is[a] = Si[a] + cn * h * di[a];
          if (ORDER >= 2) { Js[a] = Sj[a] + cn * h * dj[a]; Ts[a] = T[a] + cn * h * dd[a]; }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < NX; ++a) {
      X[a] += h / 6.0 * aX[a];
      if (ORDER >= 1) Si[a] += h / 6.0 * aI[a];
      if (ORDER >= 2) { Sj[a] += h / 6.0 * aJ[a]; T[a] += h / 6.0 * aT[a]; }
    }
  }
#endif
}

// ---- grouped linearisation: one task integrates x, up to TMPC_LIN_D first-order directions and up to TMPC_LIN_PP
// second-order pairs among them, so the ODE, its Jacobian and its second derivatives are evaluated once per group
// (TMPC_LIN_NG groups cover all NZ(NZ+1)/2 pairs; tables generated by modelgen.lin_groups).
#define TM_LD TMPC_LIN_D
#define TM_LP TMPC_LIN_PP
TM_HD constexpr int tm_g_nd(int g) { constexpr int t[] = TMPC_LIN_GND; return t[g]; }
TM_HD constexpr int tm_g_np(int g) { constexpr int t[] = TMPC_LIN_GNP; return t[g]; }
TM_HD constexpr int tm_g_dir(int g, int a) { constexpr int t[] = TMPC_LIN_GD; return t[g * TM_LD + a]; }
TM_HD constexpr int tm_g_own(int g, int a) { constexpr int t[] = TMPC_LIN_GOWN; return t[g * TM_LD + a]; }
TM_HD constexpr int tm_g_pa(int g, int p) { constexpr int t[] = TMPC_LIN_GPA; return t[g * TM_LP + p]; }
TM_HD constexpr int tm_g_pb(int g, int p) { constexpr int t[] = TMPC_LIN_GPB; return t[g * TM_LP + p]; }
TM_HD constexpr int tm_g_pi(int g, int p) { constexpr int t[] = TMPC_LIN_GPI; return t[g * TM_LP + p]; }

// derivative of the group state at stage argument (Xs, Ss, Ts).  ND directions with global ids dir[a]
template <int ND, int NP, class DirF, class PaF, class PbF>
TM_HD void tm_rhs_group(const double* Xs, const double* u, const double (*Ss)[NX], const double (*Ts)[NX], DirF dir, PaF pa,
                        PbF pb, double* k, double (*dS)[NX], double (*dT)[NX]) {
  double J[NX * NZ];
  double Hn[TMPC_NHESS > 0 ? TMPC_NHESS : 1];
  if (NP > 0) tmpc_ode_d2(Xs, u, k, J, Hn); else tmpc_ode_jac(Xs, u, k, J);
  double v[ND > 0 ? ND : 1][NZ];
#pragma unroll
  for (int a = 0; a < ND; ++a) {
#pragma unroll
    for (int i = 0; i < NX; ++i) v[a][i] = Ss[a][i];
#pragma unroll
    for (int b = 0; b < NU; ++b) v[a][NX + b] = (dir(a) == NX + b) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double t = 0.0;
#pragma unroll
      for (int b = 0; b < NZ; ++b) t += J[i * NZ + b] * v[a][b];
      dS[a][i] = t;
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    double dd[NX];
    tmpc_ode_bilin(Hn, v[pa(p)], v[pb(p)], dd);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double t = dd[i];
#pragma unroll
      for (int b = 0; b < NX; ++b) t += J[i * NZ + b] * Ts[p][b];
      dT[p][i] = t;
    }
  }
}

template <int ND, int NP, class DirF, class PaF, class PbF>
TM_HD void tm_integrate_group(const double* x0, const double* u, DirF dir, PaF pa, PbF pb, double* X, double (*S)[NX],
                              double (*T)[NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) X[i] = x0[i];
#pragma unroll
  for (int a = 0; a < ND; ++a)
#pragma unroll
    for (int i = 0; i < NX; ++i) S[a][i] = (dir(a) == i) ? 1.0 : 0.0;
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int i = 0; i < NX; ++i) T[p][i] = 0.0;
#if TMPC_DISCRETE
  {
    double k[NX], dS[ND > 0 ? ND : 1][NX], dT[NP > 0 ? NP : 1][NX];
    tm_rhs_group<ND, NP>(X, u, S, T, dir, pa, pb, k, dS, dT);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      X[i] = k[i];
#pragma unroll
      for (int a = 0; a < ND; ++a) S[a][i] = dS[a][i];
#pragma unroll
      for (int p = 0; p < NP; ++p) T[p][i] = dT[p][i];
    }
  }
#else
  const double h = TMPC_RK_DT;
  for (int s = 0; s < TMPC_RK_STEPS; ++s) {
    double aX[NX], aS[ND > 0 ? ND : 1][NX], aT[NP > 0 ? NP : 1][NX];
    double k[NX], dS[ND > 0 ? ND : 1][NX], dT[NP > 0 ? NP : 1][NX];
    double Xs[NX], Ss[ND > 0 ? ND : 1][NX], Ts[NP > 0 ? NP : 1][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      Xs[i] = X[i];
#pragma unroll
      for (int a = 0; a < ND; ++a) Ss[a][i] = S[a][i];
#pragma unroll
      for (int p = 0; p < NP; ++p) Ts[p][i] = T[p][i];
    }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      tm_rhs_group<ND, NP>(Xs, u, Ss, Ts, dir, pa, pb, k, dS, dT);
      const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
      const double cn = (st == 2) ? 1.0 : 0.5;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        aX[i] = (st == 0) ? k[i] : aX[i] + wgt * k[i];
        if (st < 3) Xs[i] = X[i] + cn * h * k[i];
#pragma unroll
        for (int a = 0; a < ND; ++a) {
          aS[a][i] = (st == 0) ? dS[a][i] : aS[a][i] + wgt * dS[a][i];
          if (st < 3) Ss[a][i] = S[a][i] + cn * h * dS[a][i];
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          aT[p][i] = (st == 0) ? dT[p][i] : aT[p][i] + wgt * dT[p][i];
          if (st < 3) Ts[p][i] = T[p][i] + cn * h * dT[p][i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      X[i] += h / 6.0 * aX[i];
#pragma unroll
      for (int a = 0; a < ND; ++a) S[a][i] += h / 6.0 * aS[a][i];
#pragma unroll
      for (int p = 0; p < NP; ++p) T[p][i] += h / 6.0 * aT[p][i];
    }
  }
#endif
}

// exact-Hessian group G: writes its owned S columns, its W entries and (group 0) xf
template <int G>
TM_HD void tm_lin_group_exact(const double* x, const double* u, const double* lam, double* rec) {
  constexpr int ND = tm_g_nd(G), NP = tm_g_np(G);
  double X[NX], S[ND][NX], T[NP][NX];
  tm_integrate_group<ND, NP>(x, u, [](int a) { return tm_g_dir(G, a); }, [](int p) { return tm_g_pa(G, p); },
                             [](int p) { return tm_g_pb(G, p); }, X, S, T);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    double wij = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) wij += lam[i] * T[p][i];
    rec[NX + NX * NZ + tm_g_pi(G, p)] = wij;
  }
#pragma unroll
  for (int a = 0; a < ND; ++a)
    if (tm_g_own(G, a)) {
#pragma unroll
      for (int i = 0; i < NX; ++i) rec[NX + i * NZ + tm_g_dir(G, a)] = S[a][i];
    }
  if (G == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) rec[i] = X[i];
  }
}

// Gauss-Newton group g: directions 3g .. 3g+2 (first order only)
#define TM_GN_NG ((NZ + 2) / 3)
template <int G>
TM_HD void tm_lin_group_gn(const double* x, const double* u, double* rec) {
  constexpr int ND = (NZ - 3 * G) < 3 ? (NZ - 3 * G) : 3;
#if TMPC_COLLOCATION
  for (int a = 0; a < ND; ++a) {                     // implicit integrator: one direction at a time
    double Xc[NX], Sc[NX], t1[NX], t2[NX];
    tm_integrate<1>(x, u, 3 * G + a, 3 * G + a, Xc, Sc, t1, t2);
    for (int i = 0; i < NX; ++i) rec[NX + i * NZ + 3 * G + a] = Sc[i];
    if (G == 0 && a == 0) for (int i = 0; i < NX; ++i) rec[i] = Xc[i];
  }
  return;
#endif
  double X[NX], S[ND][NX], T[1][NX];
  tm_integrate_group<ND, 0>(x, u, [](int a) { return 3 * G + a; }, [](int) { return 0; }, [](int) { return 0; }, X, S, T);
#pragma unroll
  for (int a = 0; a < ND; ++a)
#pragma unroll
    for (int i = 0; i < NX; ++i) rec[NX + i * NZ + 3 * G + a] = S[a][i];
  if (G == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) rec[i] = X[i];
  }
}

#if TMPC_COLLOCATION
TM_HD int tm_lin_tasks_per_stage(int) { return 1; }     // tm_colloc_stage: the whole record in one task
#elif defined(TMPC_LIN_GROUPED)
TM_HD int tm_lin_tasks_per_stage(int hessian_exact) { return hessian_exact ? TMPC_LIN_NG : TM_GN_NG; }
#else
TM_HD int tm_lin_tasks_per_stage(int hessian_exact) { return hessian_exact ? TM_NPAIR : TM_GN_NG; }
#endif

// one linearisation task.  trial = 1: evaluate at (W + D, LAMQ), else at (W, LAM).  g = group id.
TM_HD void tm_lin_task(const TmProb& P, const TmState& S, int64_t inst, int k, int g, int trial) {
  const double* w = S.W + inst * P.n_w + (int64_t)k * NZ;
  double x[NX], u[NU];
#pragma unroll
  for (int a = 0; a < NX; ++a) x[a] = w[a];
#pragma unroll
  for (int b = 0; b < NU; ++b) u[b] = w[NX + b];
  if (trial && S.qpstat[inst] != 0) return;   // failed QP: keep LIN at W for the final statistics
  if (trial) {
    const double* d = S.D + inst * P.n_w + (int64_t)k * NZ;
#pragma unroll
    for (int a = 0; a < NX; ++a) x[a] += d[a];
#pragma unroll
    for (int b = 0; b < NU; ++b) u[b] += d[NX + b];
  }
  double* rec = S.LIN + (inst * P.N + k) * (int64_t)TM_LSZ;
#if TMPC_COLLOCATION
  {
    double lamc[NX];
    const double* lamq = (trial ? S.LAMQ : S.LAM) + inst * P.n_g + tm_gdyn(P, k);
    for (int a = 0; a < NX; ++a) lamc[a] = P.hessian_exact ? lamq[a] : 0.0;
    tm_colloc_stage(x, u, P.hessian_exact ? 2 : 1, lamc, rec);
    (void)g;
    return;
  }
#endif
  if (P.hessian_exact) {
    const double* lamp = (trial ? S.LAMQ : S.LAM) + inst * P.n_g + tm_gdyn(P, k);
    double lam[NX];
#pragma unroll
    for (int a = 0; a < NX; ++a) lam[a] = lamp[a];
    // one (i,j) pair per thread: 164 registers, no spills, FP64 pipe 82 % busy.  The grouped variant
    // (tm_lin_group_exact: 3 directions + 3-4 pairs per thread, half the flops) needs 255 registers, spills, and ran
    // 3.5x slower on B200 (profiles/r01b_summary.md) -- kept for models with a cheaper right-hand side.
#ifdef TMPC_LIN_GROUPED
    switch (g) {
      case 0: tm_lin_group_exact<0>(x, u, lam, rec); break;
#if TMPC_LIN_NG > 1
      case 1: tm_lin_group_exact<1>(x, u, lam, rec); break;
#endif
#if TMPC_LIN_NG > 2
      case 2: tm_lin_group_exact<2>(x, u, lam, rec); break;
#endif
#if TMPC_LIN_NG > 3
      case 3: tm_lin_group_exact<3>(x, u, lam, rec); break;
#endif
#if TMPC_LIN_NG > 4
      case 4: tm_lin_group_exact<4>(x, u, lam, rec); break;
#endif
#if TMPC_LIN_NG > 5
      case 5: tm_lin_group_exact<5>(x, u, lam, rec); break;
#endif
#if TMPC_LIN_NG > 6
      case 6: tm_lin_group_exact<6>(x, u, lam, rec); break;
#endif
#if TMPC_LIN_NG > 7
      case 7: tm_lin_group_exact<7>(x, u, lam, rec); break;
#endif
      default: break;
    }
#else
    {
      int i, j;
      tm_pair_ij(g, i, j);
      double X[NX], Si[NX], Sj[NX], T[NX];
      tm_integrate<2>(x, u, i, j, X, Si, Sj, T);
      double wij = 0.0;
#pragma unroll
      for (int a = 0; a < NX; ++a) wij += lam[a] * T[a];
      rec[NX + NX * NZ + g] = wij;
      if (i == j) {
#pragma unroll
        for (int a = 0; a < NX; ++a) rec[NX + a * NZ + i] = Si[a];
      }
      if (g == 0) {
#pragma unroll
        for (int a = 0; a < NX; ++a) rec[a] = X[a];
      }
    }
#endif
  } else {
    switch (g) {
      case 0: tm_lin_group_gn<0>(x, u, rec); break;
#if (TMPC_NZ > 3)
      case 1: tm_lin_group_gn<1>(x, u, rec); break;
#endif
#if (TMPC_NZ > 6)
      case 2: tm_lin_group_gn<2>(x, u, rec); break;
#endif
#if (TMPC_NZ > 9)
      case 3: tm_lin_group_gn<3>(x, u, rec); break;
#endif
#if (TMPC_NZ > 12)
#error "more than 4 Gauss-Newton groups: extend the dispatch in tm_lin_task"
#endif
      default: break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: QP.  Riccati factorisation of the dynamics-constrained base problem (exact terminal penalty rho), then a
// Goldfarb-Idnani dual active-set iteration in which every other constraint (terminal equality rows first, then the
// violated inequality rows one at a time) lives in a small dense Schur complement S = N' G N, G = base inverse.
// Exact active-set solution: inactive multipliers are exact zeros, as the reference relies on (sqp_method.py:421).
// ---------------------------------------------------------------------------------------------------------------
struct TmQpWs {
  TmP AB, Q, r, b, K, Lc, hv, d, y, rhs, kk, kkm, P0, P1, PAB, F, pv, tr;
  TmP sl;                 // E = N*nh + nxt: current value of every constraint row (slack / terminal residual)
  TmP Mc;                 // M x E: column j = N G n_j of the dual Hessian for working-set member j (+ one candidate)
  TmP Lf;                 // M x M: Cholesky factor of the working-set Schur complement S = N_A G N_A'
  TmP cA, rv, nu, acts, acte, sc;   // per member: L^-1 S_Aq, S^-1 S_Aq, multiplier, sign, row id; scalars
};

TM_HD size_t tm_qpws_doubles(int N, int nh, int nxt, int M) {
  const size_t E = (size_t)N * nh + nxt;
  size_t n = 0;
  n += (size_t)N * NX * NZ;        // AB
  n += (size_t)N * NZ * NZ;        // Q
  n += (size_t)(N + 1) * NZ;       // r
  n += (size_t)N * NX;             // b
  n += (size_t)N * NU * NX;        // K
  n += (size_t)N * NU * NU;        // Lc
  n += (size_t)N * (nh > 0 ? nh : 1);   // hv
  n += 3 * (size_t)(N + 1) * NZ;   // d y rhs
  n += (size_t)N * NU;             // kk
  n += (size_t)NX * N * NU;        // kkm (feed-forward of the multi-right-hand-side terminal sweep)
  n += 2 * NX * NX + NX * NZ + NZ * NZ;   // P0 P1 PAB F
  n += 4 * NX;                     // pv (two buffers of NX, e0, spare)
  n += (nxt > 0 ? nxt : 1);        // tr
  n += (E > 0 ? E : 1);            // sl
  n += (size_t)(M + 1) * (E > 0 ? E : 1);   // Mc (+1 candidate column)
  n += (size_t)M * M;              // Lf
  n += 5 * (size_t)M + 8;          // cA rv nu acts acte sc
  return n;
}

TM_HD void tm_qpws_carve(double* base, int N, int nh, int nxt, int M, TmQpWs& s) {
  const size_t E = (size_t)N * nh + nxt;
  size_t o = 0;
#define TM_CARVE(member, n) s.member = tm_mkp(base, o); o += (size_t)(n)
  TM_CARVE(AB, (size_t)N * NX * NZ);
  TM_CARVE(Q, (size_t)N * NZ * NZ);
  TM_CARVE(r, (size_t)(N + 1) * NZ);
  TM_CARVE(b, (size_t)N * NX);
  TM_CARVE(K, (size_t)N * NU * NX);
  TM_CARVE(Lc, (size_t)N * NU * NU);
  TM_CARVE(hv, (size_t)N * (nh > 0 ? nh : 1));
  TM_CARVE(d, (size_t)(N + 1) * NZ);
  TM_CARVE(y, (size_t)(N + 1) * NZ);
  TM_CARVE(rhs, (size_t)(N + 1) * NZ);
  TM_CARVE(kk, (size_t)N * NU);
  TM_CARVE(kkm, (size_t)NX * N * NU);
  TM_CARVE(P0, NX * NX);
  TM_CARVE(P1, NX * NX);
  TM_CARVE(PAB, NX * NZ);
  TM_CARVE(F, NZ * NZ);
  TM_CARVE(pv, 4 * NX);
  TM_CARVE(tr, (nxt > 0 ? nxt : 1));
  TM_CARVE(sl, (E > 0 ? E : 1));
  TM_CARVE(Mc, (size_t)(M + 1) * (E > 0 ? E : 1));
  TM_CARVE(Lf, (size_t)M * M);
  TM_CARVE(cA, M);
  TM_CARVE(rv, M);
  TM_CARVE(nu, M);
  TM_CARVE(acts, M);
  TM_CARVE(acte, M);
  TM_CARVE(sc, 8);
#undef TM_CARVE
}

// Cholesky of the NU x NU block Fuu (row-major, in registers of every lane): returns 0 if a pivot <= thr
TM_HD int tm_chol_small(const double* Fuu, double* L, double thr) {
#pragma unroll
  for (int i = 0; i < NU * NU; ++i) L[i] = 0.0;
  for (int c = 0; c < NU; ++c) {
    double dg = Fuu[c * NU + c];
    for (int l = 0; l < c; ++l) dg -= L[c * NU + l] * L[c * NU + l];
    if (!(dg > thr)) return 0;
    const double ld = sqrt(dg);
    L[c * NU + c] = ld;
    for (int rr = c + 1; rr < NU; ++rr) {
      double v = Fuu[rr * NU + c];
      for (int l = 0; l < c; ++l) v -= L[rr * NU + l] * L[c * NU + l];
      L[rr * NU + c] = v / ld;
    }
  }
  return 1;
}
// solve (L L') x = rhs in place
template <class PL>
TM_HD void tm_chol_small_solve(PL L, double* x) {
  for (int i = 0; i < NU; ++i) {
    double v = x[i];
    for (int l = 0; l < i; ++l) v -= L[i * NU + l] * x[l];
    x[i] = v / L[i * NU + i];
  }
  for (int i = NU - 1; i >= 0; --i) {
    double v = x[i];
    for (int l = i + 1; l < NU; ++l) v -= L[l * NU + i] * x[l];
    x[i] = v / L[i * NU + i];
  }
}

// homogeneous base solve:  out = argmin 1/2 d'Qd + rhs'd  s.t. d_x0 = 0, d_x(k+1) = A d_x + B d_u   ( = -G rhs )
// kfrom: last stage with a non-zero right-hand side (N = the x_N block): the backward sweep starts there.
TM_HD void tm_ricc_solve(const TmProb& P, TmQpWs& s, TmP rhs, TmP out, int kfrom) {
  const int N = P.N;
  const int lane = TM_LANE;
  TmP pv0 = s.pv;
  TmP pv1 = s.pv + NX;
  for (int a = lane; a < NX; a += TM_NL) pv0[a] = (kfrom >= N) ? rhs[N * NZ + a] : 0.0;
  TM_SYNC();
  const int kb = (kfrom >= N) ? N - 1 : kfrom;
  for (int k = kb; k >= 0; --k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
    const TmP L = s.Lc + (size_t)k * NU * NU;
    const TmP rk = rhs + k * NZ;
    double pvr[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) pvr[i] = pv0[i];
    double fu[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double v = rk[NX + a];
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + NX + a] * pvr[i];
      fu[a] = v;
    }
    for (int j = lane; j < NX; j += TM_NL) {
      double v = rk[j];
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + j] * pvr[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) v += Kk[a * NX + j] * fu[a];
      pv1[j] = v;
    }
    double ku[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) ku[a] = -fu[a];
    tm_chol_small_solve(L, ku);
    for (int a = lane; a < NU; a += TM_NL) s.kk[k * NU + a] = ku[a];
    TM_SYNC();
    TmP t = pv0; pv0 = pv1; pv1 = t;
  }
  for (int a = lane; a < NX; a += TM_NL) out[a] = 0.0;
  TM_SYNC();
  for (int k = 0; k < N; ++k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
    const TmP dx = out + k * NZ;
    double dxr[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) dxr[j] = dx[j];
    double du[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double v = (k <= kb) ? s.kk[k * NU + a] : 0.0;
#pragma unroll
      for (int j = 0; j < NX; ++j) v += Kk[a * NX + j] * dxr[j];
      du[a] = v;
    }
    for (int a = lane; a < NU; a += TM_NL) out[k * NZ + NX + a] = du[a];
    for (int i = lane; i < NX; i += TM_NL) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < NX; ++j) v += AB[i * NZ + j] * dxr[j];
#pragma unroll
      for (int a = 0; a < NU; ++a) v += AB[i * NZ + NX + a] * du[a];
      out[(k + 1) * NZ + i] = v;
    }
    TM_SYNC();
  }
  for (int a = lane; a < NU; a += TM_NL) out[N * NZ + NX + a] = 0.0;
  TM_SYNC();
}

// ---- dual-Hessian columns ---------------------------------------------------------------------------------------
// Column for constraint row qe:  y = qs * G n_qe  is swept stage by stage and never stored; mq[e] = n_e' y for every
// row e.  The NX-wide recursions are carried in registers by every lane (too small to split; no synchronisation inside
// the sweeps); workspace traffic is the factor (AB, K, Lc), the per-stage feed-forward kk and the column itself.
TM_HD void tm_ricc_col(const TmProb& P, TmQpWs& s, int qe, double qs, TmP mq) {
  const int N = P.N, nh = P.nh, NI = N * nh;
  const int lane = TM_LANE;
  double pv[NX], rz[NZ];
#pragma unroll
  for (int a = 0; a < NX; ++a) pv[a] = 0.0;
#pragma unroll
  for (int b = 0; b < NZ; ++b) rz[b] = 0.0;
  int kb;
  if (qe >= NI) {
    const int ti = P.term_idx[qe - NI];
#pragma unroll
    for (int a = 0; a < NX; ++a) if (a == ti) pv[a] = -qs;
    kb = N - 1;
  } else {
    kb = qe / nh;
    const double* Ci = P.C + (size_t)(qe % nh) * NZ;
#pragma unroll
    for (int b = 0; b < NZ; ++b) rz[b] = -qs * Ci[b];
  }
  for (int k = kb; k >= 0; --k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
    const TmP L = s.Lc + (size_t)k * NU * NU;
    double fu[NU > 0 ? NU : 1], pn[NX];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double v = (k == kb) ? rz[NX + a] : 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + NX + a] * pv[i];
      fu[a] = v;
    }
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double v = (k == kb) ? rz[j] : 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + j] * pv[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) v += Kk[a * NX + j] * fu[a];
      pn[j] = v;
    }
    double ku[NU > 0 ? NU : 1];
#pragma unroll
    for (int a = 0; a < NU; ++a) ku[a] = -fu[a];
    tm_chol_small_solve(L, ku);
    for (int a = lane; a < NU; a += TM_NL) s.kk[k * NU + a] = ku[a];
#pragma unroll
    for (int j = 0; j < NX; ++j) pv[j] = pn[j];
  }
  TM_SYNC();
  double z[NZ];
#pragma unroll
  for (int b = 0; b < NZ; ++b) z[b] = 0.0;
  for (int k = 0; k < N; ++k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double v = (k <= kb) ? s.kk[k * NU + a] : 0.0;
#pragma unroll
      for (int j = 0; j < NX; ++j) v += Kk[a * NX + j] * z[j];
      z[NX + a] = v;
    }
    for (int i = lane; i < nh; i += TM_NL) {
      const double* Ci = P.C + (size_t)i * NZ;
      double t = 0.0;
#pragma unroll
      for (int b = 0; b < NZ; ++b) t += Ci[b] * z[b];
      mq[k * nh + i] = t;
    }
    double xn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < NZ; ++b) v += AB[i * NZ + b] * z[b];
      xn[i] = v;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) z[i] = xn[i];
  }
  for (int t = lane; t < P.nxt; t += TM_NL) {
    const int ti = P.term_idx[t];
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < NX; ++a) if (a == ti) v = z[a];
    mq[NI + t] = v;
  }
  TM_SYNC();
}

// The columns of ALL terminal rows (sign +1) in one multi-right-hand-side sweep: column t -> Mc[t*E + e].  The factor is
// read once for the nxt (<= NX) right-hand sides instead of once per row.
TM_HD void tm_ricc_cols_term(const TmProb& P, TmQpWs& s, TmP Mc, int E) {
  const int N = P.N, nh = P.nh, NI = N * nh, nxt = P.nxt;
  const int lane = TM_LANE;
  double pv[NX][NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) {
    const int ti = (c < nxt) ? P.term_idx[c] : -1;
#pragma unroll
    for (int a = 0; a < NX; ++a) pv[c][a] = (a == ti) ? -1.0 : 0.0;
  }
  for (int k = N - 1; k >= 0; --k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
    const TmP L = s.Lc + (size_t)k * NU * NU;
    double ABr[NX * NZ], Kr[NU * NX > 0 ? NU * NX : 1], Lr[NU * NU > 0 ? NU * NU : 1];
#pragma unroll
    for (int e = 0; e < NX * NZ; ++e) ABr[e] = AB[e];
#pragma unroll
    for (int e = 0; e < NU * NX; ++e) Kr[e] = Kk[e];
#pragma unroll
    for (int e = 0; e < NU * NU; ++e) Lr[e] = L[e];
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      double fu[NU > 0 ? NU : 1], pn[NX];
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) v += ABr[i * NZ + NX + a] * pv[c][i];
        fu[a] = v;
      }
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) v += ABr[i * NZ + j] * pv[c][i];
#pragma unroll
        for (int a = 0; a < NU; ++a) v += Kr[a * NX + j] * fu[a];
        pn[j] = v;
      }
      double ku[NU > 0 ? NU : 1];
#pragma unroll
      for (int a = 0; a < NU; ++a) ku[a] = -fu[a];
      tm_chol_small_solve(Lr, ku);
      for (int a = lane; a < NU; a += TM_NL) s.kkm[((size_t)c * N + k) * NU + a] = ku[a];
#pragma unroll
      for (int j = 0; j < NX; ++j) pv[c][j] = pn[j];
    }
  }
  TM_SYNC();
  double z[NX][NZ];
#pragma unroll
  for (int c = 0; c < NX; ++c)
#pragma unroll
    for (int b = 0; b < NZ; ++b) z[c][b] = 0.0;
  for (int k = 0; k < N; ++k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
    double ABr[NX * NZ], Kr[NU * NX > 0 ? NU * NX : 1];
#pragma unroll
    for (int e = 0; e < NX * NZ; ++e) ABr[e] = AB[e];
#pragma unroll
    for (int e = 0; e < NU * NX; ++e) Kr[e] = Kk[e];
#pragma unroll
    for (int c = 0; c < NX; ++c) {
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        double v = s.kkm[((size_t)c * N + k) * NU + a];
#pragma unroll
        for (int j = 0; j < NX; ++j) v += Kr[a * NX + j] * z[c][j];
        z[c][NX + a] = v;
      }
    }
    for (int i = lane; i < nh; i += TM_NL) {
      const double* Ci = P.C + (size_t)i * NZ;
#pragma unroll
      for (int c = 0; c < NX; ++c) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) t += Ci[b] * z[c][b];
        if (c < nxt) Mc[(size_t)c * E + k * nh + i] = t;
      }
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      double xn[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) v += ABr[i * NZ + b] * z[c][b];
        xn[i] = v;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) z[c][i] = xn[i];
    }
  }
  for (int t = lane; t < nxt; t += TM_NL) {
    const int ti = P.term_idx[t];
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      double v = 0.0;
#pragma unroll
      for (int a = 0; a < NX; ++a) if (a == ti) v = z[c][a];
      if (c < nxt) Mc[(size_t)c * E + NI + t] = v;
    }
  }
  TM_SYNC();
}

// constraint rows by unified id e: e < N*nh is inequality row (k = e / nh, i = e % nh), e >= N*nh terminal row
TM_HD double tm_erow_dot(const TmProb& P, int e, TmP v) {
  const int NI = P.N * P.nh;
  if (e >= NI) return v[P.N * NZ + P.term_idx[e - NI]];
  const int k = e / P.nh, i = e % P.nh;
  double t = 0.0;
  const double* Ci = P.C + (size_t)i * NZ;
  const TmP vk = v + k * NZ;
#pragma unroll
  for (int b = 0; b < NZ; ++b) t += Ci[b] * vk[b];
  return t;
}
// rhs += coef * n_e   (single lane)
TM_HD void tm_erow_axpy(const TmProb& P, int e, double coef, TmP rhs) {
  const int NI = P.N * P.nh;
  if (e >= NI) { rhs[P.N * NZ + P.term_idx[e - NI]] += coef; return; }
  const int k = e / P.nh, i = e % P.nh;
  const double* Ci = P.C + (size_t)i * NZ;
#pragma unroll
  for (int b = 0; b < NZ; ++b) rhs[k * NZ + b] += coef * Ci[b];
}

// rebuild the Cholesky factor Lf of S_ij = acts_i * Mc[j][acte_i] (i, j < m) after a deletion (single lane)
TM_HD int tm_schur_refactor(TmQpWs& s, int m, int M, int E) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j <= i; ++j) s.Lf[i * M + j] = s.acts[i] * s.Mc[(size_t)j * E + (int)s.acte[i]];
  for (int c = 0; c < m; ++c) {
    double dg = s.Lf[c * M + c];
    for (int l = 0; l < c; ++l) dg -= s.Lf[c * M + l] * s.Lf[c * M + l];
    if (!(dg > 0.0)) return 0;
    const double ld = sqrt(dg);
    s.Lf[c * M + c] = ld;
    for (int i = c + 1; i < m; ++i) {
      double v = s.Lf[i * M + c];
      for (int l = 0; l < c; ++l) v -= s.Lf[i * M + l] * s.Lf[c * M + l];
      s.Lf[i * M + c] = v / ld;
    }
  }
  return 1;
}

// returns: 0 ok, 2 infeasible / working-set overflow / numerical breakdown, 3 base factorisation not PD
// al_mask: rows (k*nh+i) whose squared slack  gamma/2 (C_i d + h_i)^2  is added to the base problem.  The term and its
// gradient vanish wherever the row is active at the QP solution, so the solution is unchanged iff every masked row ends
// up in the working set; rows that do not are reported in al_bad (caller removes them and re-solves).  This makes the
// base factorisation positive definite on the null space of (dynamics + warm-start active rows) -- the space on which
// the reference tests and regularises its reduced Hessian (sqp_method.py:335-347) -- instead of dynamics only.
// Perturbed solve used to tabulate the solution map of the first QP after reset() (tm_qp0_*, below): the equality-
// constrained part of the QP (dynamics, x_0, terminal rows; inequality rows ignored) is solved for modified data and
// the result goes to (dout, lout) instead of the instance's D / LAMQ.
struct TmQpPert {
  int homog;      // 1: zero all offsets (gradient r, dynamics defect b, terminal residual): pure linear response; the x_0 offset is always zeroed
  int e0_unit;    // >= 0: x_0 offset = unit vector e0_unit
  int row;        // >= 0: gradient -= n_row (response to a unit multiplier on inequality row `row` = k*nh + i)
  double *dout, *lout;
};

TM_HDN int tm_qp_solve(const TmProb& P, const TmState& S, int64_t inst, TmQpWs& s, int use_exact,
                       const unsigned* al_mask, unsigned* al_bad, const TmQpPert* pert = nullptr) {
  const int N = P.N, nh = P.nh, nxt = P.nxt, M = P.maxact;
  const int lane = TM_LANE;
  const double* w = S.W + inst * P.n_w;
  const double* lin = S.LIN + inst * P.N * (int64_t)TM_LSZ;
  // ---- A. stage data ------------------------------------------------------------------------------------------
  for (int e = lane; e < N * NX * NZ; e += TM_NL) { int k = e / (NX * NZ), o = e % (NX * NZ); s.AB[e] = lin[(size_t)k * TM_LSZ + NX + o]; }
  for (int e = lane; e < N * NX; e += TM_NL) { int k = e / NX, a = e % NX; s.b[e] = lin[(size_t)k * TM_LSZ + a] - w[(k + 1) * NZ + a]; }
  if (P.economic) {
    // economic stage cost: Q_k = d2l/dz2 (+ lam' d2F), r_k = dl/dz at the iterate
    for (int k = lane; k < N; k += TM_NL) {
      double z[NZ], gl[NZ], Hl[NZ * NZ];
#pragma unroll
      for (int b = 0; b < NZ; ++b) z[b] = w[k * NZ + b];
      tmpc_cost_grad(z, z + NX, gl);
      tmpc_cost_hess(z, z + NX, Hl);
      for (int i = 0; i < NZ; ++i) {
        for (int j = 0; j < NZ; ++j) {
          double v = 0.5 * (Hl[i * NZ + j] + Hl[j * NZ + i]);
          if (use_exact) v += lin[(size_t)k * TM_LSZ + NX + NX * NZ + (i <= j ? tm_pair_idx(i, j) : tm_pair_idx(j, i))];
          s.Q[(size_t)k * NZ * NZ + i * NZ + j] = v;
        }
        s.r[k * NZ + i] = gl[i];
      }
    }
  } else {
    for (int e = lane; e < N * NZ * NZ; e += TM_NL) {
      int k = e / (NZ * NZ), o = e % (NZ * NZ), i = o / NZ, j = o % NZ;
      int ph = (S.phase + k) % P.p;
      double v = P.H[(size_t)ph * NZ * NZ + o];
      if (use_exact) v += lin[(size_t)k * TM_LSZ + NX + NX * NZ + (i <= j ? tm_pair_idx(i, j) : tm_pair_idx(j, i))];
      s.Q[e] = v;
    }
    for (int e = lane; e < N * NZ; e += TM_NL) {
      int k = e / NZ, i = e % NZ;
      int ph = (S.phase + k) % P.p;
      const double* Hk = P.H + (size_t)ph * NZ * NZ + (size_t)i * NZ;
      const double* wr = P.wref + (size_t)ph * NZ;
      double v = P.q[(size_t)ph * NZ + i];
#pragma unroll
      for (int j = 0; j < NZ; ++j) v += Hk[j] * (w[k * NZ + j] - wr[j]);
      s.r[e] = v;
    }
  }
  for (int a = lane; a < NZ; a += TM_NL) s.r[N * NZ + a] = 0.0;
  for (int e = lane; e < N * nh; e += TM_NL) {
    int k = e / nh, i = e % nh;
    double v = P.c[i];
#pragma unroll
    for (int j = 0; j < NZ; ++j) v += P.C[(size_t)i * NZ + j] * w[k * NZ + j];
    s.hv[e] = v;
  }
  TmP e0 = s.pv + 2 * NX;
  for (int a = lane; a < NX; a += TM_NL) e0[a] = S.X0[inst * NX + a] - w[a];
  {
    const double* xrN = P.wref + (size_t)((S.phase + N) % P.p) * NZ;
    for (int t = lane; t < nxt; t += TM_NL) s.tr[t] = w[N * NZ + P.term_idx[t]] - xrN[P.term_idx[t]];
  }
  TM_SYNC();
  if (pert) {
    if (pert->homog) {
      for (int e = lane; e < (N + 1) * NZ; e += TM_NL) s.r[e] = 0.0;
      for (int e = lane; e < N * NX; e += TM_NL) s.b[e] = 0.0;
      for (int t = lane; t < nxt; t += TM_NL) s.tr[t] = 0.0;
    }
    for (int a = lane; a < NX; a += TM_NL) e0[a] = 0.0;     // the x_0 offset always enters through the table
    TM_SYNC();
    if (pert->e0_unit >= 0) for (int a = lane; a < NX; a += TM_NL) e0[a] = (a == pert->e0_unit) ? 1.0 : 0.0;
    if (pert->row >= 0) {
      const int k = pert->row / nh, i = pert->row % nh;
      for (int b2 = lane; b2 < NZ; b2 += TM_NL) s.r[k * NZ + b2] -= P.C[(size_t)i * NZ + b2];
    }
    TM_SYNC();
  }
  if (al_mask) {
    // gamma relative to the largest Hessian diagonal entry of the horizon
    double qmax = 0.0;
    for (int e = lane; e < N * NZ; e += TM_NL) { int k = e / NZ, i = e % NZ; qmax = fmax(qmax, fabs(s.Q[(size_t)k * NZ * NZ + i * NZ + i])); }
    qmax = tm_wmax(qmax);
    const double gam = P.al_gamma * fmax(qmax, 1e-300);
    for (int k = lane; k < N; k += TM_NL) {
      for (int i = 0; i < nh; ++i) {
        const int e = k * nh + i;
        if (!((al_mask[e >> 5] >> (e & 31)) & 1u)) continue;
        const double* Ci = P.C + (size_t)i * NZ;
        double cc = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) cc += Ci[b] * Ci[b];
        const double g = gam / cc;
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
#pragma unroll
          for (int b = 0; b < NZ; ++b) s.Q[(size_t)k * NZ * NZ + a * NZ + b] += g * Ci[a] * Ci[b];
          s.r[k * NZ + a] += g * s.hv[e] * Ci[a];
        }
      }
    }
    TM_SYNC();
  }
  // ---- B. Riccati factorisation + main solve (offsets b_k, e0, terminal penalty) -------------------------------
  TmP Pn = s.P0;   // P_{k+1}
  TmP Pk = s.P1;
  TmP pv0 = s.pv;
  TmP pv1 = s.pv + NX;
  for (int e = lane; e < NX * NX; e += TM_NL) {
    int i = e / NX, j = e % NX;
    double v = 0.0;
    for (int t = 0; t < nxt; ++t) if (P.term_idx[t] == i && i == j) v += P.rho;
    Pn[e] = v;
  }
  for (int a = lane; a < NX; a += TM_NL) {
    double v = 0.0;
    for (int t = 0; t < nxt; ++t) if (P.term_idx[t] == a) v += P.rho * s.tr[t];
    pv0[a] = v;
  }
  TM_SYNC();
  int fail = 0;
  for (int k = N - 1; k >= 0; --k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Qk = s.Q + (size_t)k * NZ * NZ;
    for (int e = lane; e < NX * NZ; e += TM_NL) {
      int i = e / NZ, c = e % NZ;
      double v = 0.0;
#pragma unroll
      for (int l = 0; l < NX; ++l) v += Pn[i * NX + l] * AB[l * NZ + c];
      s.PAB[e] = v;
    }
    TM_SYNC();
    for (int e = lane; e < NZ * NZ; e += TM_NL) {
      int a = e / NZ, c = e % NZ;
      double v = Qk[e];
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + a] * s.PAB[i * NZ + c];
      s.F[e] = v;
    }
    TM_SYNC();
    double Fuu[NU * NU], L[NU * NU];
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
      for (int c = 0; c < NU; ++c) Fuu[a * NU + c] = 0.5 * (s.F[(NX + a) * NZ + NX + c] + s.F[(NX + c) * NZ + NX + a]);
    if (!tm_chol_small(Fuu, L, P.reg_tol)) { fail = 1; break; }
    for (int e = lane; e < NU * NU; e += TM_NL) s.Lc[(size_t)k * NU * NU + e] = L[e];
    // K = -Fuu^-1 Fux : lane j owns column j
    for (int j = lane; j < NX; j += TM_NL) {
      double col[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) col[a] = -0.5 * (s.F[(NX + a) * NZ + j] + s.F[j * NZ + NX + a]);
      tm_chol_small_solve(L, col);
#pragma unroll
      for (int a = 0; a < NU; ++a) s.K[(size_t)k * NU * NX + a * NX + j] = col[a];
    }
    TM_SYNC();
    const TmP Kk = s.K + (size_t)k * NU * NX;
    // P_k = Fxx + sym(Fux' K)
    for (int e = lane; e < NX * NX; e += TM_NL) {
      int i = e / NX, j = e % NX;
      double v = 0.5 * (s.F[i * NZ + j] + s.F[j * NZ + i]);
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const double fai = 0.5 * (s.F[(NX + a) * NZ + i] + s.F[i * NZ + NX + a]);
        const double faj = 0.5 * (s.F[(NX + a) * NZ + j] + s.F[j * NZ + NX + a]);
        v += 0.5 * (fai * Kk[a * NX + j] + faj * Kk[a * NX + i]);
      }
      Pk[e] = v;
    }
    // main right-hand side: v = p_{k+1} + P_{k+1} b_k
    {
      double vv[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double t = pv0[i];
#pragma unroll
        for (int l = 0; l < NX; ++l) t += Pn[i * NX + l] * s.b[k * NX + l];
        vv[i] = t;
      }
      double fu[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        double t = s.r[k * NZ + NX + a];
#pragma unroll
        for (int i = 0; i < NX; ++i) t += AB[i * NZ + NX + a] * vv[i];
        fu[a] = t;
      }
      for (int j = lane; j < NX; j += TM_NL) {
        double t = s.r[k * NZ + j];
#pragma unroll
        for (int i = 0; i < NX; ++i) t += AB[i * NZ + j] * vv[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) t += Kk[a * NX + j] * fu[a];
        pv1[j] = t;
      }
      double ku[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) ku[a] = -fu[a];
      tm_chol_small_solve(L, ku);
      for (int a = lane; a < NU; a += TM_NL) s.kk[k * NU + a] = ku[a];
    }
    TM_SYNC();
    { TmP t = Pn; Pn = Pk; Pk = t; }
    { TmP t = pv0; pv0 = pv1; pv1 = t; }
  }
  if (fail) return 3;
  // forward sweep of the main solve
  for (int a = lane; a < NX; a += TM_NL) s.d[a] = e0[a];
  TM_SYNC();
  for (int k = 0; k < N; ++k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Kk = s.K + (size_t)k * NU * NX;
    const TmP dx = s.d + k * NZ;
    double du[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double v = s.kk[k * NU + a];
#pragma unroll
      for (int j = 0; j < NX; ++j) v += Kk[a * NX + j] * dx[j];
      du[a] = v;
    }
    for (int a = lane; a < NU; a += TM_NL) s.d[k * NZ + NX + a] = du[a];
    for (int i = lane; i < NX; i += TM_NL) {
      double v = s.b[k * NX + i];
#pragma unroll
      for (int j = 0; j < NX; ++j) v += AB[i * NZ + j] * dx[j];
#pragma unroll
      for (int a = 0; a < NU; ++a) v += AB[i * NZ + NX + a] * du[a];
      s.d[(k + 1) * NZ + i] = v;
    }
    TM_SYNC();
  }
  for (int a = lane; a < NU; a += TM_NL) s.d[N * NZ + NX + a] = 0.0;
  TM_SYNC();
  // ---- C. Goldfarb-Idnani dual active set, carried in the dual space ----------------------------------------------
  // The primal iterate is never updated inside the loop: d = d_base + sum_j nu_j G n_j is recovered by ONE solve at
  // the end.  Per added constraint: one (partial) Riccati solve y = G n_q, its dual-Hessian column Mc_q[e] = n_e'y for
  // every row e, an O(m^2) Cholesky append, and an O(E m) update of all row values sl[e].
  const int NI = N * nh, E = NI + nxt;
  for (int e = lane; e < E; e += TM_NL) s.sl[e] = (e < NI ? s.hv[e] : s.tr[e - NI]) + tm_erow_dot(P, e, s.d);
  TM_SYNC();
  int m = 0, ret = 0;
  int n_gi = 0, n_ricc = 0;
  if (nxt > 0) {
    // the terminal equality rows enter together: their columns from one multi-right-hand-side sweep, multipliers from
    // the nxt x nxt Schur complement (the state the one-at-a-time iteration would reach; equality multipliers are
    // sign-free and never dropped)
    tm_ricc_cols_term(P, s, s.Mc, E);
    n_ricc += 1;
    if (lane == 0) {
      for (int t = 0; t < nxt; ++t) { s.acte[t] = (double)(NI + t); s.acts[t] = 1.0; }
      double ok = (double)tm_schur_refactor(s, nxt, M, E);
      if (ok != 0.0) {
        for (int i = 0; i < nxt; ++i) {          // L L' nu = -sl_term
          double v = -s.sl[NI + i];
          for (int l = 0; l < i; ++l) v -= s.Lf[i * M + l] * s.nu[l];
          s.nu[i] = v / s.Lf[i * M + i];
        }
        for (int i = nxt - 1; i >= 0; --i) {
          double v = s.nu[i];
          for (int l = i + 1; l < nxt; ++l) v -= s.Lf[l * M + i] * s.nu[l];
          s.nu[i] = v / s.Lf[i * M + i];
        }
      }
      s.sc[1] = ok;
    }
    TM_SYNC();
    if (s.sc[1] == 0.0) ret = 2;
    else {
      for (int e = lane; e < E; e += TM_NL) {
        double acc = 0.0;
        for (int t = 0; t < nxt; ++t) acc += s.nu[t] * s.Mc[(size_t)t * E + e];
        s.sl[e] += acc;
      }
      m = nxt;
      TM_SYNC();
    }
  }
  const int maxit = pert ? 0 : 4 * E + 8;
  for (int it = 0; it < maxit && !ret; ++it) {
    int qe;
    double qs = 1.0, sval;
    {
      double best = TM_INF;
      int bid = 0x7fffffff;
      for (int e = lane; e < NI; e += TM_NL) {
        const int k = e / nh, i = e % nh;
        if (k == 0 && P.relax0[i]) continue;
        const double v = s.sl[e] / fmax(1.0, fabs(P.c[i]));
        if (v < best) { best = v; bid = e; }
      }
      tm_wargmin(best, bid);
      if (!(best < -1e-10)) break;            // primal feasible: optimal
      int dup = 0;                            // a working-set row can only show up here through round-off
      for (int j2 = 0; j2 < m; ++j2) if ((int)s.acte[j2] == bid) dup = 1;
      if (dup) break;
      qe = bid;
      sval = s.sl[qe];
    }
    if (m >= M) { ret = 2; break; }
    TmP mq = s.Mc + (size_t)m * E;            // candidate column, becomes member m when added
    tm_ricc_col(P, s, qe, qs, mq);
    ++n_gi; ++n_ricc;
    const double yq = qs * mq[qe];
    double nq = 0.0;
    int added = 0;
    for (int inner = 0; inner < M + 2; ++inner) {
      // l = L^-1 S_Aq (kept in cA), zn = yq - l'l, r = L^-T l (rv)
      if (lane == 0) {
        double ll = 0.0;
        for (int i = 0; i < m; ++i) {
          double v = s.acts[i] * mq[(int)s.acte[i]];
          for (int l = 0; l < i; ++l) v -= s.Lf[i * M + l] * s.cA[l];
          v /= s.Lf[i * M + i];
          s.cA[i] = v;
          ll += v * v;
        }
        for (int i = m - 1; i >= 0; --i) {
          double v = s.cA[i];
          for (int l = i + 1; l < m; ++l) v -= s.Lf[l * M + i] * s.rv[l];
          s.rv[i] = v / s.Lf[i * M + i];
        }
        s.sc[0] = ll;
      }
      TM_SYNC();
      const double zn = yq - s.sc[0];
      double t1 = TM_INF;
      int jd = -1;
      for (int j2 = 0; j2 < m; ++j2) {
        if ((int)s.acte[j2] >= NI) continue;           // equality rows are never dropped
        const double rj = s.rv[j2];
        if (rj > 1e-14) { const double tj = s.nu[j2] / rj; if (tj < t1) { t1 = tj; jd = j2; } }
      }
      const int dependent = !(zn > 1e-11 * fmax(yq, 1e-300));
      double t;
      int do_add = 0;
      if (dependent) {
        if (jd < 0) { ret = 2; break; }
        t = t1;
      } else {
        const double t2 = -sval / zn;
        if (t2 <= t1) { t = t2; do_add = 1; } else t = t1;
        for (int e = lane; e < E; e += TM_NL) {
          double acc = mq[e];
          for (int j2 = 0; j2 < m; ++j2) acc -= s.rv[j2] * s.Mc[(size_t)j2 * E + e];
          s.sl[e] += t * acc;
        }
        sval += t * zn;
      }
      TM_SYNC();
      for (int j2 = lane; j2 < m; j2 += TM_NL) s.nu[j2] -= t * s.rv[j2];
      nq += t;
      TM_SYNC();
      if (do_add) {
        if (lane == 0) {
          for (int l = 0; l < m; ++l) s.Lf[m * M + l] = s.cA[l];
          s.Lf[m * M + m] = sqrt(zn);
          s.acte[m] = (double)qe; s.acts[m] = qs; s.nu[m] = nq;
        }
        TM_SYNC();
        ++m;
        added = 1;
        break;
      }
      // drop member jd: shift members jd+1..m-1 and the candidate column down by one slot, rebuild the factor
      for (int a = jd; a < m; ++a) {
        for (int e = lane; e < E; e += TM_NL) s.Mc[(size_t)a * E + e] = s.Mc[(size_t)(a + 1) * E + e];
        TM_SYNC();
      }
      if (lane == 0) {
        for (int a = jd; a < m - 1; ++a) { s.acte[a] = s.acte[a + 1]; s.acts[a] = s.acts[a + 1]; s.nu[a] = s.nu[a + 1]; }
      }
      --m;
      TM_SYNC();
      mq = s.Mc + (size_t)m * E;
      if (lane == 0) s.sc[1] = (double)tm_schur_refactor(s, m, M, E);
      TM_SYNC();
      if (s.sc[1] == 0.0) { ret = 2; break; }
    }
    if (ret) break;
    if (!added) { ret = 2; break; }
    if (it == maxit - 1) ret = 2;
  }
  if (!ret) {
    // primal recovery: d = d_base + G N_A' nu  (one full solve)
    for (int e = lane; e < (N + 1) * NZ; e += TM_NL) s.rhs[e] = 0.0;
    TM_SYNC();
    int kfrom = -1;
    if (lane == 0) for (int j2 = 0; j2 < m; ++j2) tm_erow_axpy(P, (int)s.acte[j2], -s.acts[j2] * s.nu[j2], s.rhs);
    for (int j2 = 0; j2 < m; ++j2) { const int e = (int)s.acte[j2]; const int k = e < NI ? e / nh : N; if (k > kfrom) kfrom = k; }
    TM_SYNC();
    if (m > 0) {
      tm_ricc_solve(P, s, s.rhs, s.y, kfrom);
      ++n_ricc;
      for (int e = lane; e < (N + 1) * NZ; e += TM_NL) s.d[e] += s.y[e];
      TM_SYNC();
    }
  }
  if (lane == 0 && !pert) {
    S.qpwork[inst] = n_gi;
#ifdef __CUDA_ARCH__
    atomicAdd(S.counters + 5, 1ull); atomicAdd(S.counters + 6, (unsigned long long)n_gi); atomicAdd(S.counters + 7, (unsigned long long)n_ricc);
#else
    S.counters[5] += 1; S.counters[6] += n_gi; S.counters[7] += n_ricc;
#endif
  }
  if (ret) return ret;
  if (al_mask) {
    int nbad = 0;
    for (int wd = 0; wd < TM_ALW; ++wd) al_bad[wd] = 0u;
    for (int e = 0; e < NI && e < 32 * TM_ALW; ++e) {
      if (!((al_mask[e >> 5] >> (e & 31)) & 1u)) continue;
      int found = 0;
      for (int j2 = 0; j2 < m; ++j2) if ((int)s.acte[j2] == e) found = 1;
      if (!found) { al_bad[e >> 5] |= (1u << (e & 31)); ++nbad; }
    }
    if (nbad) return 5;
  }
  // ---- D. outputs: step and multipliers (CasADi sign: H d + g + J' lam = 0) ------------------------------------
  double* dout = pert ? pert->dout : S.D + inst * P.n_w;
  double* lq = pert ? pert->lout : S.LAMQ + inst * P.n_g;
  for (int e = lane; e < P.n_w; e += TM_NL) dout[e] = s.d[e];
  for (int e = lane; e < P.n_g; e += TM_NL) lq[e] = 0.0;
  TM_SYNC();
  if (lane == 0) {
    for (int j2 = 0; j2 < m; ++j2) {
      const int e = (int)s.acte[j2];
      if (e >= NI) lq[tm_gterm(P) + (e - NI)] = -s.acts[j2] * s.nu[j2];
      else lq[tm_gh(P, e / nh) + e % nh] = -s.nu[j2];
    }
  }
  TM_SYNC();
  // dynamics multipliers by the stationarity recursion (pv0/pv1 reused as lam_{k}, lam_{k-1})
  pv0 = s.pv; pv1 = s.pv + NX;
  for (int a = lane; a < NX; a += TM_NL) {
    double v = 0.0;
    for (int t = 0; t < nxt; ++t) if (P.term_idx[t] == a) v += lq[tm_gterm(P) + t];
    pv0[a] = v;
  }
  TM_SYNC();
  for (int k = N - 1; k >= 0; --k) {
    const TmP AB = s.AB + (size_t)k * NX * NZ;
    const TmP Qk = s.Q + (size_t)k * NZ * NZ;
    for (int a = lane; a < NX; a += TM_NL) lq[tm_gdyn(P, k) + a] = pv0[a];
    for (int j = lane; j < NX; j += TM_NL) {
      double v = s.r[k * NZ + j];
#pragma unroll
      for (int c = 0; c < NZ; ++c) v += 0.5 * (Qk[j * NZ + c] + Qk[c * NZ + j]) * s.d[k * NZ + c];
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + j] * pv0[i];
      for (int i = 0; i < nh; ++i) v += P.C[(size_t)i * NZ + j] * lq[tm_gh(P, k) + i];
      pv1[j] = v;
    }
    TM_SYNC();
    { TmP t = pv0; pv0 = pv1; pv1 = t; }
  }
  for (int a = lane; a < NX; a += TM_NL) lq[a] = -pv0[a];
  TM_SYNC();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// objective and infeasibility of a point w_t = W + alpha*D.  xf_from_lin: take F(x_k,u_k) from LIN (valid for the
// point LIN was evaluated at), else integrate (lane <-> stage).  Writes g (n_g) when gout != nullptr.
// ---------------------------------------------------------------------------------------------------------------
TM_HD void tm_eval_point(const TmProb& P, const TmState& S, int64_t inst, double alpha, int xf_from_lin,
                         double* gout, double& f_out, double& viol_out) {
  const int N = P.N, nh = P.nh;
  const int lane = TM_LANE;
  const double* w = S.W + inst * P.n_w;
  const double* d = S.D + inst * P.n_w;
  const double* lin = S.LIN + inst * P.N * (int64_t)TM_LSZ;
  double f = 0.0, viol = 0.0;
  for (int k = lane; k < N; k += TM_NL) {
    double z[NZ], xn[NX], xf[NX];
#pragma unroll
    for (int b = 0; b < NZ; ++b) z[b] = w[k * NZ + b] + (alpha != 0.0 ? alpha * d[k * NZ + b] : 0.0);
#pragma unroll
    for (int a = 0; a < NX; ++a) xn[a] = w[(k + 1) * NZ + a] + (alpha != 0.0 ? alpha * d[(k + 1) * NZ + a] : 0.0);
    if (xf_from_lin) {
#pragma unroll
      for (int a = 0; a < NX; ++a) xf[a] = lin[(size_t)k * TM_LSZ + a];
    } else {
      double t1[NX], t2[NX], t3[NX];
      tm_integrate<0>(z, z + NX, 0, 0, xf, t1, t2, t3);
    }
    const int ph = (S.phase + k) % P.p;
    const double* Hk = P.H + (size_t)ph * NZ * NZ;
    const double* wr = P.wref + (size_t)ph * NZ;
    const double* qk = P.q + (size_t)ph * NZ;
    double dz[NZ];
#pragma unroll
    for (int b = 0; b < NZ; ++b) dz[b] = z[b] - wr[b];
    double fk = 0.0;
    if (P.economic) {
      fk = tmpc_stage_cost(z, z + NX);
    } else {
#pragma unroll
      for (int a = 0; a < NZ; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) t += Hk[a * NZ + b] * dz[b];
        fk += dz[a] * (0.5 * t + qk[a]);
      }
    }
    f += fk;
#pragma unroll
    for (int a = 0; a < NX; ++a) {
      const double rdy = xf[a] - xn[a];
      viol = fmax(viol, fabs(rdy));
      if (gout) gout[tm_gdyn(P, k) + a] = rdy;
    }
    for (int i = 0; i < nh; ++i) {
      double v = P.c[i];
#pragma unroll
      for (int b = 0; b < NZ; ++b) v += P.C[(size_t)i * NZ + b] * z[b];
      if (gout) gout[tm_gh(P, k) + i] = v;
      if (!(k == 0 && P.relax0[i]) && v < 0.0) viol = fmax(viol, -v);
    }
    if (k == 0) {
#pragma unroll
      for (int a = 0; a < NX; ++a) {
        const double r0 = z[a] - S.X0[inst * NX + a];
        viol = fmax(viol, fabs(r0));
        if (gout) gout[a] = r0;
      }
    }
    if (k == N - 1) {
      const double* xrN = P.wref + (size_t)((S.phase + N) % P.p) * NZ;
      for (int t = 0; t < P.nxt; ++t) {
        const double rt = xn[P.term_idx[t]] - xrN[P.term_idx[t]];
        viol = fmax(viol, fabs(rt));
        if (gout) gout[tm_gterm(P) + t] = rt;
      }
    }
  }
  f_out = tm_wsum(f);
  viol_out = tm_wmax(viol);
}

// |grad_w L|_inf at (W, LAM) using the linearisation stored in LIN (must be valid at W)      sqp_method.py:246
TM_HD double tm_dual_infeas(const TmProb& P, const TmState& S, int64_t inst) {
  const int N = P.N, nh = P.nh;
  const int lane = TM_LANE;
  const double* w = S.W + inst * P.n_w;
  const double* lam = S.LAM + inst * P.n_g;
  const double* lin = S.LIN + inst * P.N * (int64_t)TM_LSZ;
  double mx = 0.0;
  for (int k = lane; k < N; k += TM_NL) {
    const int ph = (S.phase + k) % P.p;
    const double* Hk = P.H + (size_t)ph * NZ * NZ;
    const double* wr = P.wref + (size_t)ph * NZ;
    const double* qk = P.q + (size_t)ph * NZ;
    const double* AB = lin + (size_t)k * TM_LSZ + NX;
    const double* ld = lam + tm_gdyn(P, k);
    const double* lh = lam + tm_gh(P, k);
    double dz[NZ], gl[NZ];
#pragma unroll
    for (int b = 0; b < NZ; ++b) dz[b] = w[k * NZ + b] - wr[b];
    if (P.economic) {
      double z[NZ];
#pragma unroll
      for (int b = 0; b < NZ; ++b) z[b] = w[k * NZ + b];
      tmpc_cost_grad(z, z + NX, gl);
    }
#pragma unroll
    for (int a = 0; a < NZ; ++a) {
      double v = P.economic ? gl[a] : qk[a];
      if (!P.economic) {
#pragma unroll
        for (int b = 0; b < NZ; ++b) v += Hk[a * NZ + b] * dz[b];
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) v += AB[i * NZ + a] * ld[i];
      for (int i = 0; i < nh; ++i) v += P.C[(size_t)i * NZ + a] * lh[i];
      if (a < NX) v += (k == 0) ? lam[a] : -lam[tm_gdyn(P, k - 1) + a];
      mx = fmax(mx, fabs(v));
    }
    if (k == N - 1) {
#pragma unroll
      for (int a = 0; a < NX; ++a) {
        double v = -ld[a];
        for (int t = 0; t < P.nxt; ++t) if (P.term_idx[t] == a) v += lam[tm_gterm(P) + t];
        mx = fmax(mx, fabs(v));
      }
    }
  }
  return tm_wmax(mx);
}

TM_HD int tm_is_ineq_active(const TmProb& P, const double* lam, int e) {   // e = k*nh + i
  return lam[tm_gh(P, e / P.nh) + e % P.nh] != 0.0;
}

// k = 0 bookkeeping (sqp_method.py:248-261): filter <- [(f0, infeas0)], as_idx_init.  LIN valid at (W, LAM).
TM_HD void tm_init(const TmProb& P, const TmState& S, int64_t inst) {
  double f, v;
  tm_eval_point(P, S, inst, 0.0, 1, nullptr, f, v);
  const int lane = TM_LANE;
  if (lane == 0) {
    S.FILT[inst * P.filter_cap * 2 + 0] = f;
    S.FILT[inst * P.filter_cap * 2 + 1] = v;
    S.nfilt[inst] = 1;
    S.iter[inst] = 0;
    S.status[inst] = -1;
    S.flags[inst] = 0;
    S.qpmode[inst] = 0;
    S.qpstat[inst] = 0;
    const double* lam = S.LAM + inst * P.n_g;
    for (int wd = 0; wd < S.aswords; ++wd) {
      unsigned bits = 0;
      for (int bt = 0; bt < 32; ++bt) {
        int e = wd * 32 + bt;
        if (e < P.N * P.nh && tm_is_ineq_active(P, lam, e)) bits |= (1u << bt);
      }
      S.asinit[inst * S.aswords + wd] = bits;
    }
  }
  TM_SYNC();
}

// finalise an instance: stats of __postprocessing (sqp_method.py:203-219) and __detect_AC (pmpc.py:840-856)
TM_HD void tm_finalize(const TmProb& P, const TmState& S, int64_t inst, int status) {
  double f, v;
  tm_eval_point(P, S, inst, 0.0, 1, S.G + inst * P.n_g, f, v);
  if (TM_LANE == 0) {
    const double* lam = S.LAM + inst * P.n_g;
    int nAS = 0, nACt = 0, nAC0 = 0;
    for (int e = 0; e < P.N * P.nh; ++e) {
      const int a = tm_is_ineq_active(P, lam, e);
      const int a0 = (S.asinit[inst * S.aswords + e / 32] >> (e % 32)) & 1u;
      nAS += a;
      nACt += (a != a0);
    }
    const double* lref = P.ref_du + (size_t)S.phase * P.n_g;
    for (int i = 0; i < P.nh; ++i) nAC0 += ((lam[tm_gh(P, 0) + i] != 0.0) != (lref[tm_gh(P, 0) + i] != 0.0));
    S.fval[inst] = f;
    S.nAS[inst] = nAS;
    S.nACtot[inst] = nACt;
    S.nAC[inst] = nAC0;
    S.status[inst] = status;
  }
  TM_SYNC();
}

// convergence test at the accepted point (LIN valid at W, LAM):  sqp_method.py:276-283
TM_HD void tm_conv(const TmProb& P, const TmState& S, int64_t inst) {
  const double dual = tm_dual_infeas(P, S, inst);
  const int nf = S.nfilt[inst];
  const double viol = S.FILT[(inst * P.filter_cap + nf - 1) * 2 + 1];
  const int it = S.iter[inst];
  int done = -1;
  if (!(dual == dual) || !(viol == viol)) done = 4;
  else if (viol < P.tol && dual < P.tol) done = 0;
  else if (it >= P.max_iter || nf >= P.filter_cap) done = 1;
  if (done >= 0) {
    tm_finalize(P, S, inst, done);
  } else if (TM_LANE == 0) {
#ifdef __CUDA_ARCH__
    int pos = atomicAdd(S.cnt_next, 1);
#else
    int pos = (*S.cnt_next)++;
#endif
    S.list_next[pos] = (int)inst;
  }
  TM_SYNC();
}

// after the QP and the trial linearisation at (W + D, LAMQ): filter line search, update, convergence
TM_HD void tm_post(const TmProb& P, const TmState& S, int64_t inst) {
  const int lane = TM_LANE;
  const int qp_status = S.qpstat[inst];
  if (qp_status != 0) { tm_finalize(P, S, inst, qp_status); return; }
  double alpha = 1.0, f, v;
  tm_eval_point(P, S, inst, alpha, 1, nullptr, f, v);
  const int nf = S.nfilt[inst];
  const double* F = S.FILT + inst * P.filter_cap * 2;
  int relin = 0;
  unsigned long long ndyn = 0;
  for (int ls = 0; ls < P.max_ls; ++ls) {                       // sqp_method.py:304-320
    int cnt = 0;
    for (int e = lane; e < nf; e += TM_NL) cnt += (f > F[2 * e] && v > F[2 * e + 1]) ? 1 : 0;
    cnt = tm_wsumi(cnt);
    if (cnt > 1) {
      alpha *= P.beta;
      tm_eval_point(P, S, inst, alpha, 0, nullptr, f, v);
      relin = 1;
      ndyn += (unsigned long long)P.N;
    } else break;
  }
  if (!(f == f) || !(v == v)) { tm_finalize(P, S, inst, 4); return; }
  double* w = S.W + inst * P.n_w;
  const double* d = S.D + inst * P.n_w;
  double* lam = S.LAM + inst * P.n_g;
  const double* lq = S.LAMQ + inst * P.n_g;
  for (int e = lane; e < P.n_w; e += TM_NL) w[e] += alpha * d[e];        // :174
  for (int e = lane; e < P.n_g; e += TM_NL) lam[e] = lq[e];              // :175 full dual step
  if (lane == 0) {
    S.FILT[(inst * P.filter_cap + nf) * 2 + 0] = f;                       // :323
    S.FILT[(inst * P.filter_cap + nf) * 2 + 1] = v;
    S.nfilt[inst] = nf + 1;
    S.iter[inst] += 1;
    if (relin) S.flags[inst] |= 2;
#ifdef __CUDA_ARCH__
    atomicAdd(S.counters + 0, 1ull);
    if (ndyn) atomicAdd(S.counters + 4, ndyn);
#else
    S.counters[0] += 1; S.counters[4] += ndyn;
#endif
  }
  TM_SYNC();
  if (relin) {
    if (lane == 0) {
#ifdef __CUDA_ARCH__
      int pos = atomicAdd(S.cnt_relin, 1);
#else
      int pos = (*S.cnt_relin)++;
#endif
      S.list_relin[pos] = (int)inst;
    }
    TM_SYNC();
    return;
  }
  tm_conv(P, S, inst);
}

// warm-start shift (pmpc.py:867-906): (W, LAM) -> (Ws, Ls); one warp per instance
TM_HD void tm_shift(const TmProb& P, const double* w, const double* lam, double* ws, double* ls) {
  const int N = P.N, nh = P.nh;
  const int lane = TM_LANE;
  for (int e = lane; e < P.n_w; e += TM_NL) {
    int k = e / NZ, o = e % NZ;
    double v;
    if (k >= N) v = w[N * NZ + o];                                   // x_N <- x_N
    else if (o < NX) v = w[(k + 1) * NZ + o];                        // x_i <- x_{i+1}  (x_{N-1} <- x_N)
    else v = (k < N - 1) ? w[(k + 1) * NZ + o] : w[(N - 1) * NZ + o];  // u_{N-1} <- u_{N-2}^{shifted} = u_{N-1}
    ws[e] = v;
  }
  for (int e = lane; e < P.n_g; e += TM_NL) {
    double v;
    if (e < NX) v = lam[tm_gdyn(P, 0) + e];                          // init <- dyn_0
    else if (e >= tm_gterm(P)) v = lam[e];                           // term kept
    else {
      int k = (e - NX) / (NX + nh), o = (e - NX) % (NX + nh);
      int ksrc = (k < N - 1) ? k + 1 : N - 1;                        // last stage duplicated
      v = lam[NX + ksrc * (NX + nh) + o];
    }
    ls[e] = v;
  }
  TM_SYNC();
}

// __prefilter_lam_g (sqp_method.py:223-238): zero every multiplier below lam_tresh, equality rows included
TM_HD void tm_prefilter(const TmProb& P, const TmState& S, int64_t inst) {
  double* lam = S.LAM + inst * P.n_g;
  for (int e = TM_LANE; e < P.n_g; e += TM_NL) if (fabs(lam[e]) < P.lam_tresh) lam[e] = 0.0;
  TM_SYNC();
}

// ---------------------------------------------------------------------------------------------------------------
// First QP after reset(): every instance starts from the same (w0, lam0) (pmpc.py:930-942), so all B QPs share the
// Hessian, the constraint Jacobian and every offset except the x_0 residual e0 = x0 - w0[0:nx]: ONE parametric QP.
// Its equality-constrained solution map is tabulated once (tm_qp_solve with TmQpPert, one warp per table row):
//     TAB[0]          (d, lam) for e0 = 0                 TAB[1+a]       response to e0 = unit_a
//     TAB[1+NX+e]     response to a unit multiplier on inequality row e (terminal rows stay enforced)
//     SL0 / SLPHI     row values n_e'd of TAB[0] (+ h(w0)) / of TAB[1+a];     MCOL[e][e'] = n_e' G' n_e
// and each instance runs only the Goldfarb-Idnani working-set iteration on these tables (thread per instance, state in
// registers / local memory, no Riccati sweep, no per-instance workspace), then one table combination for (d, lam).
// Same exact active-set solution as tm_qp_solve (unique: strictly convex on the feasible null space).
// ---------------------------------------------------------------------------------------------------------------
#define TM_Q0_MAXM 20      /* working-set capacity (inequality rows); overflow -> generic path */
struct TmQp0Tab {
  int EI, EIs, n_out, nT;  // inequality rows N*nh, padded row stride of MCOL (odd), n_w + n_g, 1 + NX + EI
  double *TAB, *SL0, *SLPHI, *MCOL;
  int* bad;                // != 0: the tabulation failed (base factorisation not PD ...) -> every instance takes the generic path
};

// returns 0 ok / 2 not solved here (overflow, dependent rows, breakdown): the caller queues the instance for tm_qp
TM_HD int tm_qp0_gi(const TmProb& P, const TmQp0Tab& T, const double* e0, int* acte, double* nu, int& m_out, int& ngi_out) {
  const int EI = T.EI, EIs = T.EIs, nh = P.nh;
  double Lf[TM_Q0_MAXM * (TM_Q0_MAXM + 1) / 2], cA[TM_Q0_MAXM], rv[TM_Q0_MAXM];   // Lf: packed lower triangle
#define TM_LF(i, j) Lf[(i) * ((i) + 1) / 2 + (j)]
  int m = 0, ngi = 0;
  m_out = 0; ngi_out = 0;
  const int maxit = 4 * EI + 8;
  for (int it = 0; it < maxit; ++it) {
    double best = TM_INF, bval = 0.0;
    int bid = -1;
    for (int e = 0; e < EI; ++e) {
      const int k = e / nh, i = e - k * nh;
      if (k == 0 && P.relax0[i]) continue;
      double v = T.SL0[e];
#pragma unroll
      for (int a = 0; a < NX; ++a) v += e0[a] * T.SLPHI[a * EI + e];
      for (int j = 0; j < m; ++j) v += nu[j] * T.MCOL[(size_t)acte[j] * EIs + e];
      const double sc = v / fmax(1.0, fabs(P.c[i]));
      if (sc < best) { best = sc; bid = e; bval = v; }
    }
    if (!(best < -1e-10)) break;                 // primal feasible: optimal
    int dup = 0;
    for (int j = 0; j < m; ++j) if (acte[j] == bid) dup = 1;
    if (dup) break;
    if (m >= TM_Q0_MAXM) return 2;
    const int qe = bid;
    double sval = bval;
    const double yq = T.MCOL[(size_t)qe * EIs + qe];
    double nq = 0.0;
    int added = 0;
    ++ngi;
    for (int inner = 0; inner < TM_Q0_MAXM + 2; ++inner) {
      double ll = 0.0;
      for (int i = 0; i < m; ++i) {
        double v = T.MCOL[(size_t)qe * EIs + acte[i]];
        for (int l = 0; l < i; ++l) v -= TM_LF(i, l) * cA[l];
        v /= TM_LF(i, i);
        cA[i] = v;
        ll += v * v;
      }
      for (int i = m - 1; i >= 0; --i) {
        double v = cA[i];
        for (int l = i + 1; l < m; ++l) v -= TM_LF(l, i) * rv[l];
        rv[i] = v / TM_LF(i, i);
      }
      const double zn = yq - ll;
      double t1 = TM_INF;
      int jd = -1;
      for (int j = 0; j < m; ++j) {
        if (rv[j] > 1e-14) { const double tj = nu[j] / rv[j]; if (tj < t1) { t1 = tj; jd = j; } }
      }
      const int dependent = !(zn > 1e-11 * fmax(yq, 1e-300));
      double t;
      int do_add = 0;
      if (dependent) {
        if (jd < 0) return 2;
        t = t1;
      } else {
        const double t2 = -sval / zn;
        if (t2 <= t1) { t = t2; do_add = 1; } else t = t1;
        sval += t * zn;
      }
      for (int j = 0; j < m; ++j) nu[j] -= t * rv[j];
      nq += t;
      if (do_add) {
        for (int l = 0; l < m; ++l) TM_LF(m, l) = cA[l];
        TM_LF(m, m) = sqrt(zn);
        acte[m] = qe; nu[m] = nq;
        ++m;
        added = 1;
        break;
      }
      for (int a = jd; a < m - 1; ++a) { acte[a] = acte[a + 1]; nu[a] = nu[a + 1]; }
      --m;
      for (int i = 0; i < m; ++i)
        for (int j = 0; j <= i; ++j) TM_LF(i, j) = T.MCOL[(size_t)acte[j] * EIs + acte[i]];
      for (int c = 0; c < m; ++c) {
        double dg = TM_LF(c, c);
        for (int l = 0; l < c; ++l) dg -= TM_LF(c, l) * TM_LF(c, l);
        if (!(dg > 0.0)) return 2;
        const double ld = sqrt(dg);
        TM_LF(c, c) = ld;
        for (int i = c + 1; i < m; ++i) {
          double v = TM_LF(i, c);
          for (int l = 0; l < c; ++l) v -= TM_LF(i, l) * TM_LF(c, l);
          TM_LF(i, c) = v / ld;
        }
      }
    }
    if (!added) return 2;
    if (it == maxit - 1) return 2;
  }
  m_out = m; ngi_out = ngi;
  return 0;
}
#undef TM_LF

// (d, lam)[i] of one instance from the tables: element i of  TAB[0] + sum_a e0_a TAB[1+a] + sum_j nu_j TAB[1+NX+acte_j]
TM_HD double tm_qp0_combine(const TmQp0Tab& T, int i, const double* e0, const int* acte, const double* nu, int m) {
  double v = T.TAB[i];
#pragma unroll
  for (int a = 0; a < NX; ++a) v += e0[a] * T.TAB[(size_t)(1 + a) * T.n_out + i];
  for (int j = 0; j < m; ++j) v += nu[j] * T.TAB[(size_t)(1 + NX + acte[j]) * T.n_out + i];
  return v;
}

// table row t of the tabulation (one warp / one twin call per row); inst = any instance (all identical)
TM_HD void tm_qp0_build_row(const TmProb& P, const TmState& S, int64_t inst, TmQpWs& ws, const TmQp0Tab& T, int t) {
  TmQpPert pt;
  pt.homog = t > 0;
  pt.e0_unit = (t >= 1 && t <= NX) ? t - 1 : -1;
  pt.row = t > NX ? t - 1 - NX : -1;
  pt.dout = T.TAB + (size_t)t * T.n_out;
  pt.lout = pt.dout + P.n_w;
  unsigned bad[TM_ALW];
  const int ret = tm_qp_solve(P, S, inst, ws, P.hessian_exact, nullptr, bad, &pt);
  if (TM_LANE == 0) {
    if (ret != 0) {
#ifdef __CUDA_ARCH__
      atomicExch(T.bad, 1);
#else
      *T.bad = 1;
#endif
    } else if (pt.row >= 0) {
      pt.lout[tm_gh(P, pt.row / P.nh) + pt.row % P.nh] = -1.0;       // the unit multiplier itself (lam_h = -nu)
    }
  }
  TM_SYNC();
}

// derived tables: entry (t, e) = n_e' TAB[t].d  (+ h_e(w0) for t = 0)
TM_HD void tm_qp0_derive(const TmProb& P, const TmState& S, int64_t inst, const TmQp0Tab& T, int t, int e) {
  const int k = e / P.nh, i = e % P.nh;
  const double* Ci = P.C + (size_t)i * NZ;
  const double* d = T.TAB + (size_t)t * T.n_out + (size_t)k * NZ;
  double v = 0.0;
#pragma unroll
  for (int b = 0; b < NZ; ++b) v += Ci[b] * d[b];
  if (t == 0) {
    const double* w = S.W + inst * P.n_w + (size_t)k * NZ;
    double hv = P.c[i];
#pragma unroll
    for (int b = 0; b < NZ; ++b) hv += Ci[b] * w[b];
    T.SL0[e] = hv + v;
  } else if (t <= NX) {
    T.SLPHI[(size_t)(t - 1) * T.EI + e] = v;
  } else {
    T.MCOL[(size_t)(t - 1 - NX) * T.EIs + e] = v;
  }
}

// bookkeeping of one instance solved (ret == 0) or not (ret != 0 -> queued for the generic path) by the table route
TM_HD void tm_qp0_finish(const TmProb& P, const TmState& S, int64_t inst, int ret, int ngi) {
  if (ret == 0) {
    S.qpmode[inst] = 0;
    S.qpstat[inst] = 0;
    S.qpwork[inst] = ngi;
#ifdef __CUDA_ARCH__
    atomicAdd(S.counters + 5, 1ull); atomicAdd(S.counters + 6, (unsigned long long)ngi); atomicAdd(S.counters + 8, 1ull);
#else
    S.counters[5] += 1; S.counters[6] += ngi; S.counters[8] += 1;
#endif
  } else {
    S.qpmode[inst] = 0;
#ifdef __CUDA_ARCH__
    const int pos = atomicAdd(S.cnt_retry, 1);
#else
    const int pos = (*S.cnt_retry)++;
#endif
    S.list_retry[pos] = (int)inst;
  }
}

// One QP attempt with the configured Hessian.  Exact mode: (1) augmented-Lagrangian convexification on the rows that
// are active in the current multipliers (the reference's reduced space); if some of those rows turn out inactive they
// are removed from the mask and the instance is queued for a re-solve; (2) if the base factorisation is still not
// positive definite: re-solve with the Gauss-Newton Hessian, flagged (the reference would eigen-clip its reduced
// Hessian there, sqp_method.py:345-376).  Re-solves run in a later launch over the compacted retry list, so that the
// lanes of a warp stay in step (thread-per-instance kernel).
TM_HD void tm_qp(const TmProb& P, const TmState& S, int64_t inst, TmQpWs& ws) {
  unsigned mask[TM_ALW], bad[TM_ALW];
  int nmask = 0;
  const int mode = S.qpmode[inst];
  const int use_exact = P.hessian_exact && mode < 100;
  for (int wd = 0; wd < TM_ALW; ++wd) mask[wd] = 0u;
  if (use_exact && P.al_gamma > 0.0) {
    if (mode == 0) {
      const double* lam = S.LAM + inst * P.n_g;
      for (int e = 0; e < P.N * P.nh && e < 32 * TM_ALW; ++e) {
        const int k = e / P.nh, i = e % P.nh;
        if (k == 0 && P.relax0[i]) continue;
        if (lam[tm_gh(P, k) + i] != 0.0) { mask[e >> 5] |= (1u << (e & 31)); ++nmask; }
      }
    } else {
      for (int wd = 0; wd < TM_ALW; ++wd) { mask[wd] = S.almask[inst * TM_ALW + wd]; nmask += (mask[wd] != 0u); }
    }
  }
  int ret = tm_qp_solve(P, S, inst, ws, use_exact, nmask ? mask : nullptr, bad);
  if (TM_LANE == 0) {   // attempt histogram: [8 + 4*(0 fresh | 1 mask retry | 2 Gauss-Newton) + (0 ok | 1 infeasible | 2 not PD | 3 mask rows inactive)]
    const int hm = mode == 0 ? 0 : (mode < 100 ? 1 : 2), hr = ret == 0 ? 0 : (ret == 2 ? 1 : (ret == 3 ? 2 : 3));
#ifdef __CUDA_ARCH__
    atomicAdd(S.counters + 8 + 4 * hm + hr, 1ull);
#else
    S.counters[8 + 4 * hm + hr] += 1;
#endif
  }
  int next_mode = 0;
  if (ret == 5) next_mode = (mode + 1 >= 4) ? 100 : mode + 1;
  else if (ret == 3 && use_exact) next_mode = 100;
  if (TM_LANE == 0) {
    if (next_mode) {
      for (int wd = 0; wd < TM_ALW; ++wd) S.almask[inst * TM_ALW + wd] = (ret == 5) ? (mask[wd] & ~bad[wd]) : 0u;
      S.qpmode[inst] = next_mode;
      if (next_mode == 100) S.flags[inst] |= 1;
#ifdef __CUDA_ARCH__
      const int pos = atomicAdd(S.cnt_retry, 1);
#else
      const int pos = (*S.cnt_retry)++;
#endif
      S.list_retry[pos] = (int)inst;
    } else {
      S.qpmode[inst] = 0;
      S.qpstat[inst] = ret;
    }
  }
  TM_SYNC();
}

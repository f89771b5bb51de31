// tmpc_lin2.cuh -- K1 (stage linearisation) as a warp-specialised kernel: the production linearisation for RK4 models.
//
// Replaces in the reference: g_fun / jacg_fun / H_fun evaluated by CasADi AD through the integrator
// (tunempc/sqp_method.py:152,159,330; map over stages tunempc/pmpc.py:262-266).
//
// Why this shape (profiles/r01a_summary.md): with one thread per (instance, stage, pair) every one of the NZ(NZ+1)/2 pair
// threads re-evaluates the ODE, its Jacobian, its second derivatives and both exp() at all 4*M RK4 stage points -- two
// thirds of the FP64 issue slots went to that redundant common part.  The common part depends on the state trajectory
// only, never on the sensitivities.  Here
//     lane  <-> task (instance, stage)           : 32 tasks per CTA, every lane of a warp runs the same code
//     warp  <-> role                              : warp 0 = producer (integrates x, evaluates f, J, d2f at each stage
//                                                   point ONCE per task and publishes J / d2f in shared memory),
//                                                   warps 1.. = consumers (own <= TMPC_L2_ND first-order directions
//                                                   and <= TMPC_L2_NP second-order pairs each, tables from modelgen)
// The producer runs one stage evaluation ahead of the consumers (double-buffered exchange, one __syncthreads per
// stage evaluation).  Stage arguments of the first-order columns S_i are exchanged between consumers through shared
// memory as well, so a pair (i,j) can live in any warp.  All state (x, S, T) stays in registers; HBM traffic is the
// task's (x,u,lam) in and the 49-double record out.
#pragma once
#include "tmpc_core.cuh"

#if TMPC_RK4

#define L2_NCW TMPC_L2_NCW
#define L2_THREADS (32 * (1 + L2_NCW))
#define L2_NC (NX * NZ + (TMPC_NHESS > 0 ? TMPC_NHESS : 1))   /* doubles per task in one exchange buffer: J | d2f */
#define L2_NV (NZ * NX)                                       /* stage arguments of all first-order columns */

__host__ __device__ constexpr int l2_jnz(int e) { constexpr int t[] = TMPC_JNZ; return t[e]; }
#define L2_NP TMPC_L2_NP
// role tables (modelgen.lin2_roles): every consumer warp runs the SAME code on a star of pairs (centre, partner_p)
__constant__ int l2_cen[L2_NCW] = TMPC_L2_CEN;
__constant__ int l2_own[L2_NCW] = TMPC_L2_OWN;
__constant__ int l2_part[L2_NCW * L2_NP] = TMPC_L2_PART;

static inline size_t tm_lin2_smem_bytes() { return (size_t)(2 * L2_NC + 2 * L2_NV) * 32 * sizeof(double); }

// ---- producer: x trajectory + the common evaluations -------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ void l2_producer(const double* x0, const double* u, double* Cb, int lane, double* Xout) {
  double X[NX], Xs[NX], aX[NX];
#pragma unroll
  for (int a = 0; a < NX; ++a) { X[a] = x0[a]; Xs[a] = x0[a]; aX[a] = 0.0; }
  const double h = TMPC_RK_DT;
#pragma unroll 1
  for (int e = 0; e < 4 * TMPC_RK_STEPS; ++e) {
    const int st = e & 3;
    double k[NX], J[NX * NZ], Hn[TMPC_NHESS > 0 ? TMPC_NHESS : 1];
    if (EXACT) tmpc_ode_d2(Xs, u, k, J, Hn); else tmpc_ode_jac(Xs, u, k, J);
    double* cb = Cb + (size_t)(e & 1) * L2_NC * 32 + lane;
#pragma unroll
    for (int i = 0; i < NX * NZ; ++i) if (l2_jnz(i)) cb[i * 32] = J[i];
    if (EXACT) {
#pragma unroll
      for (int i = 0; i < TMPC_NHESS; ++i) cb[(NX * NZ + i) * 32] = Hn[i];
    }
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double cnh = ((st == 2) ? 1.0 : 0.5) * h;
#pragma unroll
    for (int a = 0; a < NX; ++a) {
      aX[a] += wgt * k[a];
      Xs[a] = X[a] + cnh * k[a];
    }
    if (st == 3) {
#pragma unroll
      for (int a = 0; a < NX; ++a) { X[a] += h / 6.0 * aX[a]; Xs[a] = X[a]; aX[a] = 0.0; }
    }
    __syncthreads();
  }
  __syncthreads();   // the consumers' last stage evaluation
#pragma unroll
  for (int a = 0; a < NX; ++a) Xout[a] = X[a];
}

// ---- consumer (generic role) ---------------------------------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ void l2_consumer(int role, const double* lam, const double* Cb, double* Vb, int lane,
                                            double* rec, bool valid) {
  constexpr int NP = EXACT ? L2_NP : 0;
  constexpr int NPa = NP > 0 ? NP : 1;
  const int cen = l2_cen[role];
  const bool own = l2_own[role] != 0;
  int part[NPa];
#pragma unroll
  for (int p = 0; p < NP; ++p) part[p] = l2_part[role * L2_NP + p];
  double S[NX], aS[NX];
  double T[NPa][NX], Ts[NPa][NX], aT[NPa][NX];
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    S[r] = (cen == r) ? 1.0 : 0.0;
    aS[r] = 0.0;
    if (own) Vb[(size_t)(cen * NX + r) * 32 + lane] = S[r];   // buffer 0: stage argument of stage evaluation 0
  }
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int r = 0; r < NX; ++r) { T[p][r] = 0.0; Ts[p][r] = 0.0; aT[p][r] = 0.0; }
  double uc[NU > 0 ? NU : 1];   // input part of the centre direction (unit vector for input directions)
#pragma unroll
  for (int b = 0; b < NU; ++b) uc[b] = (cen == NX + b) ? 1.0 : 0.0;
  __syncthreads();   // iteration 0: the producer fills exchange buffer 0
  const double h = TMPC_RK_DT;
#pragma unroll 1
  for (int s4 = 0; s4 < TMPC_RK_STEPS; ++s4) {
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const double* cb = Cb + (size_t)(st & 1) * L2_NC * 32 + lane;
    const double* vb = Vb + (size_t)(st & 1) * L2_NV * 32 + lane;
    double* vn = Vb + (size_t)((st + 1) & 1) * L2_NV * 32 + lane;
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double cnh = ((st == 2) ? 1.0 : 0.5) * h;
    double J[NX * NZ];
#pragma unroll
    for (int i = 0; i < NX * NZ; ++i) J[i] = l2_jnz(i) ? cb[i * 32] : 0.0;
    double vc[NZ];
#pragma unroll
    for (int r = 0; r < NX; ++r) vc[r] = vb[(cen * NX + r) * 32];
#pragma unroll
    for (int b = 0; b < NU; ++b) vc[NX + b] = uc[b];
    // first-order column of the centre direction: dS = J * [Ss ; e_u]
    if (own) {
#pragma unroll
      for (int r = 0; r < NX; ++r) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) if (l2_jnz(r * NZ + b)) t += J[r * NZ + b] * vc[b];
        if (st == 0) aS[r] = t; else aS[r] += wgt * t;
        if (st < 3) vn[(cen * NX + r) * 32] = S[r] + cnh * t;
        else { S[r] += h / 6.0 * aS[r]; vn[(cen * NX + r) * 32] = S[r]; }
      }
    }
    // second-order pairs (centre, partner):  dT = d2f(v_c, v_p) + Jx * Ts,  d2f(v_c, .) contracted once per stage
    if (NP > 0) {
      double G[TMPC_NG];
      {
        double Hn[TMPC_NHESS > 0 ? TMPC_NHESS : 1];
#pragma unroll
        for (int i = 0; i < TMPC_NHESS; ++i) Hn[i] = cb[(NX * NZ + i) * 32];
        tmpc_ode_hv(Hn, vc, G);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int dj = part[p] < 0 ? cen : part[p];       // unused slot: harmless duplicate, never written out
        double vj[NZ], dd[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) vj[r] = vb[(dj * NX + r) * 32];
#pragma unroll
        for (int b = 0; b < NU; ++b) vj[NX + b] = (dj == NX + b) ? 1.0 : 0.0;
        tmpc_ode_gw(G, vj, dd);
#pragma unroll
        for (int r = 0; r < NX; ++r) {
          double t = dd[r];
#pragma unroll
          for (int b = 0; b < NX; ++b) if (l2_jnz(r * NZ + b)) t += J[r * NZ + b] * Ts[p][b];
          dd[r] = t;
        }
#pragma unroll
        for (int r = 0; r < NX; ++r) {
          if (st == 0) aT[p][r] = dd[r]; else aT[p][r] += wgt * dd[r];
          if (st < 3) Ts[p][r] = T[p][r] + cnh * dd[r];
          else { T[p][r] += h / 6.0 * aT[p][r]; Ts[p][r] = T[p][r]; }
        }
      }
    }
    __syncthreads();
  }
  }
  if (!valid) return;
  if (own) {
#pragma unroll
    for (int r = 0; r < NX; ++r) rec[NX + r * NZ + cen] = S[r];
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (part[p] < 0) continue;
    double wij = 0.0;
#pragma unroll
    for (int r = 0; r < NX; ++r) wij += lam[r] * T[p][r];
    const int i = cen < part[p] ? cen : part[p], j = cen < part[p] ? part[p] : cen;
    rec[NX + NX * NZ + tm_pair_idx(i, j)] = wij;
  }
}

// one CTA = 32 tasks (slot*N + k); trial = 1: evaluate at (W + D, LAMQ), else at (W, LAM)
#ifndef TMPC_L2_MINB
#define TMPC_L2_MINB 1
#endif
template <bool EXACT>
__global__ void __launch_bounds__(L2_THREADS, TMPC_L2_MINB) k_lin2(TmProb P, TmState S, const int* list, const int* cnt_dev, int cnt,
                                                     int trial) {
  extern __shared__ double l2_smem[];
  double* Cb = l2_smem;
  double* Vb = l2_smem + 2 * L2_NC * 32;
  if (cnt_dev) cnt = *cnt_dev;
  const int lane = threadIdx.x & 31;
  const int role = (int)(threadIdx.x >> 5) - 1;      // -1 producer
  const int64_t total = (int64_t)cnt * P.N;
  const int64_t first = (int64_t)blockIdx.x * 32;
  if (first >= total) return;                        // whole CTA leaves together
  int64_t t = first + lane;
  bool valid = t < total;
  if (!valid) t = total - 1;                         // idle lanes shadow the last task (no writes)
  const int k = (int)(t % P.N);
  const int64_t slot = t / P.N;
  const int64_t inst = list ? list[slot] : slot;
  if (trial && S.qpstat[inst] != 0) valid = false;   // failed QP: keep LIN at W for the final statistics
  double* rec = S.LIN + (inst * P.N + k) * (int64_t)TM_LSZ;
  if (role < 0) {
    const double* w = S.W + inst * P.n_w + (int64_t)k * NZ;
    double x[NX], u[NU], xf[NX];
#pragma unroll
    for (int a = 0; a < NX; ++a) x[a] = w[a];
#pragma unroll
    for (int b = 0; b < NU; ++b) u[b] = w[NX + b];
    if (trial) {
      const double* d = S.D + inst * P.n_w + (int64_t)k * NZ;
#pragma unroll
      for (int a = 0; a < NX; ++a) x[a] += d[a];
#pragma unroll
      for (int b = 0; b < NU; ++b) u[b] += d[NX + b];
    }
    l2_producer<EXACT>(x, u, Cb, lane, xf);
    if (valid) {
#pragma unroll
      for (int a = 0; a < NX; ++a) rec[a] = xf[a];
    }
  } else {
    double lam[NX];
    const double* lamp = (trial ? S.LAMQ : S.LAM) + inst * P.n_g + tm_gdyn(P, k);
#pragma unroll
    for (int a = 0; a < NX; ++a) lam[a] = EXACT ? lamp[a] : 0.0;
    l2_consumer<EXACT>(role, lam, Cb, Vb, lane, rec, valid);
  }
}

#endif  // TMPC_RK4

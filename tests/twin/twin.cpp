// CPU twin of the device routines -- TEST INFRASTRUCTURE ONLY (never loaded by the product package).
// Compiles tunempc_b200/csrc/tmpc_core.cuh with g++ as a 1-lane sequential program (TM_NL = 1) and runs the same
// SQP loop the CUDA host code runs, instance by instance.  Used by the CPU-only tests to check the algorithm
// (stage linearisation, Riccati + dual active set QP, filter line search, convergence, shift) against the oracle
// without a GPU; on the GPU box the same comparison runs against the real kernels.
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <vector>
#include "tmpc_core.cuh"

extern "C" {

void twin_model_info(int* nx, int* nu) { *nx = NX; *nu = NUM; }

// stage evaluation through the pair-wise integrator (same code path as tmpc_stage_eval_host)
void twin_stage_eval(int n, const double* x, const double* u, int order, double* xf, double* S, double* T) {
  for (int s = 0; s < n; ++s) {
    const double* xs = x + (size_t)s * NX;
    const double* us = u + (size_t)s * NUM;
    if (order == 0) {
      double t1[NX], t2[NX], t3[NX];
      tm_integrate<0>(xs, us, 0, 0, xf + (size_t)s * NX, t1, t2, t3);
      continue;
    }
#pragma push_macro("NZ")
#undef NZ
#define NZ TMPC_NZM
    for (int i = 0; i < NZ; ++i)
      for (int j = i; j < NZ; ++j) {
        if (order == 1 && j != i) continue;
        double X[NX], Si[NX], Sj[NX], Tt[NX];
        if (order == 1) tm_integrate<1>(xs, us, i, j, X, Si, Sj, Tt);
        else tm_integrate<2>(xs, us, i, j, X, Si, Sj, Tt);
        for (int a = 0; a < NX; ++a) {
          xf[(size_t)s * NX + a] = X[a];
          if (i == j) S[((size_t)s * NX + a) * NZ + i] = Si[a];
          if (order == 2) {
            T[(((size_t)s * NX + a) * NZ + i) * NZ + j] = Tt[a];
            T[(((size_t)s * NX + a) * NZ + j) * NZ + i] = Tt[a];
          }
        }
      }
#pragma pop_macro("NZ")
  }
}

// forward / adjoint stage linearisation (tmpc_lin3.cuh): record xf | S | W for n stage points with multipliers lam (n x nx)
int twin_lin_adjoint(int n, const double* x, const double* u, const double* lam, int order, double* rec) {
#if TMPC_RK4
  for (int s = 0; s < n; ++s) tm_lin_adjoint(x + (size_t)s * NX, u + (size_t)s * NUM, order, lam + (size_t)s * NX, rec + (size_t)s * TM_LSZ);
  return 0;
#else
  (void)n; (void)x; (void)u; (void)lam; (void)order; (void)rec;
  return 1;
#endif
}

// dims: N nh nxt p ; iopts: hessian_exact max_iter max_ls maxact economic ; dopts: tol lam_tresh beta reg_tol rho_rel
int twin_step(const int* dims, const int* iopts, const double* dopts, const double* wref, const double* H,
              const double* q, const double* ref_du, const double* C, const double* c, const int* term_idx,
              const int* relax0, int phase, long long B, const double* X0, double* W, double* LAM, double* G,
              int* status, int* iter, int* flags, double* fval, int* nAS, int* nACtot, int* nAC, double* Wsh,
              double* Lsh, long long* counters_out, int uniform) {
  TmProb P;
  P.N = dims[0]; P.nh = dims[1]; P.nxt = dims[2]; P.p = dims[3];
  P.n_w = P.N * NZ + NX;
  P.n_g = NX + P.N * (NX + NS + P.nh) + P.nxt;
  P.hessian_exact = iopts[0];
  P.filter_cap = 64;
  P.max_iter = iopts[1];
  P.max_ls = iopts[2];
  P.maxact = (iopts[3] < P.N * P.nh ? iopts[3] : P.N * P.nh) + P.nxt; if (P.maxact < 1) P.maxact = 1;
  P.economic = iopts[4];
  P.reg_mode = iopts[5];
  P.nonconvex_after = iopts[6];
  P.lin_adjoint = (getenv("TWIN_LIN_ADJOINT") && TMPC_RK4) ? 1 : 0;
  if (P.economic) P.hessian_exact = 1;
  P.tol = dopts[0]; P.lam_tresh = dopts[1]; P.beta = dopts[2]; P.reg_tol = dopts[3]; P.rho_rel = dopts[4];
  P.wref = wref; P.H = H; P.q = q; P.ref_du = ref_du; P.C = C; P.c = c; P.term_idx = term_idx; P.relax0 = relax0;
  std::vector<int> rowpin((size_t)(P.nh > 0 ? P.nh : 1), -1);
  for (int i = 0; i < P.nh; ++i) {
    int cnt = 0, jc = -1;
    for (int cidx = 0; cidx < NZ; ++cidx) if (C[(size_t)i * NZ + cidx] != 0.0) { ++cnt; jc = cidx; }
    if (cnt == 1 && jc >= NX) rowpin[i] = jc - NX;
  }
  P.rowpin = rowpin.data();
  P.prof_counters = nullptr;
  TmState S;
  memset(&S, 0, sizeof S);
  S.B = B; S.phase = phase; S.X0 = X0; S.W = W; S.LAM = LAM; S.G = G;
  S.status = status; S.iter = iter; S.flags = flags; S.fval = fval; S.nAS = nAS; S.nACtot = nACtot; S.nAC = nAC;
  std::vector<double> D((size_t)B * P.n_w), LQ((size_t)B * P.n_g), LIN((size_t)B * P.N * TM_LSZ),
      FILT((size_t)B * P.filter_cap * 2);
  std::vector<int> nfilt(B), qpstat(B, 0), qpmode(B, 0), qpwork(B, 0), la(B), lb(B), lrel(B), lretry(B);
  std::vector<unsigned> almask((size_t)B * TM_ALW);
  int cnt_retry = 0;
  S.aswords = (P.N * P.nh + 31) / 32; if (S.aswords < 1) S.aswords = 1;
  std::vector<unsigned> asinit((size_t)B * S.aswords);
  unsigned long long counters[TM_NCNT] = {0};
  int cnts[2] = {0, 0};
  S.D = D.data(); S.LAMQ = LQ.data(); S.LIN = LIN.data(); S.FILT = FILT.data(); S.nfilt = nfilt.data();
  S.qpstat = qpstat.data(); S.qpmode = qpmode.data(); S.qpwork = qpwork.data(); S.almask = almask.data(); S.list_retry = lretry.data(); S.cnt_retry = &cnt_retry; S.asinit = asinit.data(); S.counters = counters;
  S.cnt_next = &cnts[0]; S.cnt_relin = &cnts[1]; S.list_relin = lrel.data();
  const int per = tm_lin_tasks_per_stage(P.hessian_exact);
  std::vector<double> wsbuf(tm_qpws_doubles(P.N, P.nh, P.nxt, P.maxact));
  TmQpWs ws;
  tm_qpws_carve(wsbuf.data(), P.N, P.nh, P.nxt, P.maxact, ws);

  for (long long i = 0; i < B; ++i) tm_prefilter(P, S, i);
  for (long long i = 0; i < B; ++i) for (int k = 0; k < P.N; ++k) for (int pr = 0; pr < per; ++pr) tm_lin_task(P, S, i, k, pr, 0);
  for (long long i = 0; i < B; ++i) tm_init(P, S, i);
  std::vector<int>* cur = &la; std::vector<int>* nxt = &lb;
  long long nact = B;
  for (long long i = 0; i < B; ++i) la[i] = (int)i;
  long long nqp = 0, nlin = B * P.N;
  int guard = 0;
  while (nact > 0) {
    S.list_next = nxt->data();
    cnts[0] = cnts[1] = 0;
    cnt_retry = 0;
    if (guard == 0 && uniform && P.nh > 0) {
      // first QP after reset: the shared-table route of the CUDA host loop (k_qp0_build / k_qp0_derive / k_qp0)
      TmQp0Tab T;
      T.EI = P.N * P.nh; T.EIs = T.EI | 1; T.n_out = P.n_w + P.n_g; T.nT = 1 + NX + T.EI;
      std::vector<double> TAB((size_t)T.nT * T.n_out), SL0(T.EI), SLPHI((size_t)NX * T.EI), MCOL((size_t)T.EI * T.EIs, 0.0);
      int bad = 0;
      T.TAB = TAB.data(); T.SL0 = SL0.data(); T.SLPHI = SLPHI.data(); T.MCOL = MCOL.data(); T.bad = &bad;
      for (int t = 0; t < T.nT; ++t) tm_qp0_build_row(P, S, 0, ws, T, t);
      for (int t = 0; t < T.nT; ++t) for (int e = 0; e < T.EI; ++e) tm_qp0_derive(P, S, 0, T, t, e);
      if (getenv("TWIN_DUMP")) {
        FILE* f = fopen(getenv("TWIN_DUMP"), "wb");
        int hdr[4] = {T.EI, T.EIs, T.n_out, T.nT};
        fwrite(hdr, sizeof(int), 4, f);
        fwrite(TAB.data(), sizeof(double), TAB.size(), f); fwrite(SL0.data(), sizeof(double), SL0.size(), f);
        fwrite(SLPHI.data(), sizeof(double), SLPHI.size(), f); fwrite(MCOL.data(), sizeof(double), MCOL.size(), f);
        fclose(f);
      }
      for (long long s = 0; s < nact; ++s) {
        const long long i = (*cur)[s];
        if (bad) { tm_qp0_finish(P, S, i, 2, 0); continue; }
        double e0[NX], nu[TM_Q0_MAXM];
        int acte[TM_Q0_MAXM], m = 0, ngi = 0;
        for (int a = 0; a < NX; ++a) e0[a] = X0[i * NX + a] - W[i * P.n_w + a];
        const int ret = tm_qp0_gi(P, T, e0, acte, nu, m, ngi);
        tm_qp0_finish(P, S, i, ret, ngi);
        if (ret) continue;
        for (int o = 0; o < T.n_out; ++o) {
          const double v = tm_qp0_combine(T, o, e0, acte, nu, m);
          if (o < P.n_w) D[(size_t)i * P.n_w + o] = v; else LQ[(size_t)i * P.n_g + o - P.n_w] = v;
        }
      }
    } else
    for (long long s = 0; s < nact; ++s) tm_qp(P, S, (*cur)[s], ws);
    if (cnt_retry > 0) {                                         // instances the shared-table route handed back
      std::vector<int> todo(lretry.begin(), lretry.begin() + cnt_retry);
      cnt_retry = 0;
      for (int v : todo) tm_qp(P, S, v, ws);
    }
    for (long long s = 0; s < nact; ++s) for (int k = 0; k < P.N; ++k) for (int pr = 0; pr < per; ++pr) tm_lin_task(P, S, (*cur)[s], k, pr, 1);
    for (long long s = 0; s < nact; ++s) tm_post(P, S, (*cur)[s]);
    if (getenv("TWIN_TRACE")) {
      for (long long s = 0; s < nact; ++s) {
        const int i = (*cur)[s];
        double dn = 0.0; for (int e = 0; e < P.n_w; ++e) dn = fmax(dn, fabs(D[(size_t)i * P.n_w + e]));
        int nact_l = 0; for (int e = 0; e < P.N * P.nh; ++e) nact_l += tm_is_ineq_active(P, LAM + (size_t)i * P.n_g, e);
        fprintf(stderr, "[twin] it %d inst %d flags %d qpstat %d |d| %.3e nAS %d f %.9e viol %.3e attempts %llu gn %llu\n", guard, i, flags[i], qpstat[i], dn, nact_l,
                FILT[((size_t)i * P.filter_cap + nfilt[i] - 1) * 2], FILT[((size_t)i * P.filter_cap + nfilt[i] - 1) * 2 + 1], counters[5], counters[16]);
      }
    }
    nqp += nact; nlin += nact * P.N;
    const int nrel = cnts[1];
    for (int s = 0; s < nrel; ++s) for (int k = 0; k < P.N; ++k) for (int pr = 0; pr < per; ++pr) tm_lin_task(P, S, lrel[s], k, pr, 0);
    for (int s = 0; s < nrel; ++s) tm_conv(P, S, lrel[s]);
    nlin += (long long)nrel * P.N;
    std::swap(cur, nxt);
    nact = cnts[0];
    if (++guard > P.max_iter + 2) return 9;
  }
  for (long long i = 0; i < B; ++i) tm_pd_check(P, S, i, ws);          // sqp_method.py:190-201
  for (long long i = 0; i < B; ++i) tm_shift(P, W + i * P.n_w, LAM + i * P.n_g, Wsh + i * P.n_w, Lsh + i * P.n_g);
  if (counters_out) { counters_out[0] = (long long)counters[0]; counters_out[2] = nqp; counters_out[3] = nlin; counters_out[4] = (long long)counters[4]; counters_out[5] = (long long)counters[5]; counters_out[6] = (long long)counters[6]; counters_out[7] = (long long)counters[7]; for (int i = 8; i < TM_NCNT; ++i) counters_out[i] = (long long)counters[i]; }
  return 0;
}

}  // extern "C"

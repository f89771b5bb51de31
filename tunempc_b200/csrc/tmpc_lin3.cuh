// tmpc_lin3.cuh -- K1 (stage linearisation) by a forward / adjoint sweep: one task = one (instance, stage), one thread.
//
// Replaces in the reference: g_fun / jacg_fun / H_fun evaluated by CasADi AD through the integrator
// (tunempc/sqp_method.py:152,159,330; map over stages tunempc/pmpc.py:262-266).
//
// What the QP needs from a stage is  xf = F(x,u),  S = dF/dz (nx x nz)  and  W = d2(lam'F)/dz2 (nz x nz, symmetric) -- ONE
// scalar function's Hessian, not the nx x nz x nz tensor.  The pair-wise kernels (k_lin, k_lin2) propagate all nz(nz+1)/2
// second-order direction pairs forwards and contract with lam at the end.  Here the RK4 map x_{n+1} = Phi(x_n, u) is
// differentiated the other way round (value-function recursion of phi_n(x_n,u) = lam' x_M):
//   forward :  x_0 .. x_{M-1} stored (M*nx doubles of local memory)
//   backward:  for n = M-1 .. 0, with l = d phi_{n+1}/dx (nx), Psi = d2 phi_{n+1}/d(x,u)2 (packed), R = dx_M/d(x_{n+1},u):
//                stage points X_i, J_i = df/dz(X_i), adjoint weights mu_i of the stage derivatives k_i (from l),
//                V_i = dX_i/d(x_n,u) forwards over the 4 stages,  G = sum_i [V_i;E]' (mu_i . d2f(X_i)) [V_i;E],
//                [A B] = dPhi/d(x_n,u),   Psi <- [A B;0 I]' Psi [A B;0 I] + G,   l <- A'l,   R <- R [A B;0 I]
//   result  :  xf = x_M,  S = R,  W = Psi.
// Exact derivatives of the same discrete RK4 map (agrees with the pair-wise route to round-off), at ~1/3 of its FP64
// instruction count: nothing is propagated per direction pair.
#pragma once

#if TMPC_RK4

TM_HD constexpr int tm_jnz(int e) { constexpr int t[] = TMPC_JNZ; return t[e]; }
TM_HD constexpr int tm_hess_a(int e) { constexpr int t[] = TMPC_HESS_A; return t[e]; }
TM_HD constexpr int tm_hess_b(int e) { constexpr int t[] = TMPC_HESS_B; return t[e]; }
TM_HD constexpr int tm_hess_c(int e) { constexpr int t[] = TMPC_HESS_C; return t[e]; }

// G (packed i<=j) += [Vx;E]' M [Vx;E]  with M = sum_a mu_a d2f_a/dz2 given by the structural non-zeros Hn, Vx = NX x NZ,
// E = [0 I_nu]  (rows of the stage argument (X,u) w.r.t. the directions (x_n,u))
template <bool FIRST>
TM_HD void tm_adj_hess_acc(const double* Hn, const double* mu, const double* Vx, double* G) {
  // y = M [Vx;E]  (NZ x NZ), built from the non-zeros
  double y[NZ * NZ];
#pragma unroll
  for (int e = 0; e < NZ * NZ; ++e) y[e] = 0.0;
#pragma unroll
  for (int i = 0; i < TMPC_NHESS; ++i) {
    const int a = tm_hess_a(i), b = tm_hess_b(i), c = tm_hess_c(i);
    const double m = mu[a] * Hn[i];
#pragma unroll
    for (int q = 0; q < NZ; ++q) {
      // row b of y gets m * (row c of [Vx;E]); and symmetrically
      const double vc = (c < NX) ? (FIRST ? (c == q ? 1.0 : 0.0) : Vx[c * NZ + q]) : (q == c ? 1.0 : 0.0);
      y[b * NZ + q] += m * vc;
      if (b != c) {
        const double vb = (b < NX) ? (FIRST ? (b == q ? 1.0 : 0.0) : Vx[b * NZ + q]) : (q == b ? 1.0 : 0.0);
        y[c * NZ + q] += m * vb;
      }
    }
  }
  // G[r,q] += sum_b [Vx;E][b,r] y[b,q],  r <= q
#pragma unroll
  for (int r = 0; r < NZ; ++r)
#pragma unroll
    for (int q = r; q < NZ; ++q) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < NX; ++b) s += (FIRST ? (b == r ? 1.0 : 0.0) : Vx[b * NZ + r]) * y[b * NZ + q];
      if (r >= NX) s += y[r * NZ + q];
      G[tm_pair_idx(r, q)] += s;
    }
}

// dk = Jx Vx + [0 Ju]   (NX x NZ);  FIRST: Vx = [I 0]
template <bool FIRST>
TM_HD void tm_adj_dk(const double* J, const double* Vx, double* dk) {
#pragma unroll
  for (int a = 0; a < NX; ++a)
#pragma unroll
    for (int q = 0; q < NZ; ++q) {
      double s = (q >= NX && tm_jnz(a * NZ + q)) ? J[a * NZ + q] : 0.0;
#pragma unroll
      for (int b = 0; b < NX; ++b)
        if (tm_jnz(a * NZ + b)) s += J[a * NZ + b] * (FIRST ? (b == q ? 1.0 : 0.0) : Vx[b * NZ + q]);
      dk[a * NZ + q] = s;
    }
}

// order 1: xf, S;  order 2: + W = d2(lam'F)/dz2.  rec layout as everywhere: xf | S row-major | W packed.
TM_HD void tm_lin_adjoint(const double* x0, const double* u, int order, const double* lam, double* rec) {
  const double h = TMPC_RK_DT;
  double Xc[TMPC_RK_STEPS][NX];
  double X[NX];
#pragma unroll
  for (int a = 0; a < NX; ++a) X[a] = x0[a];
#pragma unroll 1
  for (int n = 0; n < TMPC_RK_STEPS; ++n) {
    double k[NX], Xs[NX], aX[NX];
#pragma unroll
    for (int a = 0; a < NX; ++a) { Xc[n][a] = X[a]; Xs[a] = X[a]; }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      tmpc_ode(Xs, u, k);
      const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
      const double cn = (st == 2) ? 1.0 : 0.5;
#pragma unroll
      for (int a = 0; a < NX; ++a) {
        aX[a] = (st == 0) ? k[a] : aX[a] + wgt * k[a];
        if (st < 3) Xs[a] = X[a] + cn * h * k[a];
      }
    }
#pragma unroll
    for (int a = 0; a < NX; ++a) X[a] += h / 6.0 * aX[a];
  }
#pragma unroll
  for (int a = 0; a < NX; ++a) rec[a] = X[a];

  double l[NX], R[NX * NZ], Psi[TM_NPAIR];
#pragma unroll
  for (int a = 0; a < NX; ++a) l[a] = (order == 2) ? lam[a] : 0.0;
#pragma unroll
  for (int e = 0; e < NX * NZ; ++e) R[e] = (e / NZ == e % NZ) ? 1.0 : 0.0;
#pragma unroll
  for (int e = 0; e < TM_NPAIR; ++e) Psi[e] = 0.0;

#pragma unroll 1
  for (int n = TMPC_RK_STEPS - 1; n >= 0; --n) {
    double xn[NX];
#pragma unroll
    for (int a = 0; a < NX; ++a) xn[a] = Xc[n][a];
    // stage points, Jacobians, second derivatives
    double J[4][NX * NZ], Hn[4][TMPC_NHESS > 0 ? TMPC_NHESS : 1];
    {
      double Xs[NX], k[NX];
#pragma unroll
      for (int a = 0; a < NX; ++a) Xs[a] = xn[a];
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        if (order == 2) tmpc_ode_d2(Xs, u, k, J[st], Hn[st]); else tmpc_ode_jac(Xs, u, k, J[st]);
        const double cn = (st == 2) ? 1.0 : 0.5;
        if (st < 3) {
#pragma unroll
          for (int a = 0; a < NX; ++a) Xs[a] = xn[a] + cn * h * k[a];
        }
      }
    }
    // adjoint weights of the stage derivatives:  mu4 = h/6 l, mu3 = h/3 l + h J4x' mu4, mu2 = h/3 l + h/2 J3x' mu3, mu1 = h/6 l + h/2 J2x' mu2
    double mu[4][NX];
    if (order == 2) {
#pragma unroll
      for (int a = 0; a < NX; ++a) mu[3][a] = h / 6.0 * l[a];
#pragma unroll
      for (int st = 2; st >= 0; --st) {
        const double cf = (st == 2) ? h : 0.5 * h;
        const double wl = (st == 0) ? h / 6.0 : h / 3.0;
#pragma unroll
        for (int b = 0; b < NX; ++b) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < NX; ++a) if (tm_jnz(a * NZ + b)) s += J[st + 1][a * NZ + b] * mu[st + 1][a];
          mu[st][b] = wl * l[b] + cf * s;
        }
      }
    }
    // forwards over the stages: V_i, G, [A B]
    double G[TM_NPAIR], Vx[NX * NZ], AB[NX * NZ], dk[NX * NZ];
#pragma unroll
    for (int e = 0; e < TM_NPAIR; ++e) G[e] = 0.0;
    if (order == 2) tm_adj_hess_acc<true>(Hn[0], mu[0], Vx, G);
    tm_adj_dk<true>(J[0], Vx, dk);
#pragma unroll
    for (int e = 0; e < NX * NZ; ++e) { AB[e] = dk[e]; Vx[e] = ((e / NZ == e % NZ) ? 1.0 : 0.0) + 0.5 * h * dk[e]; }
#pragma unroll
    for (int st = 1; st < 4; ++st) {
      if (order == 2) tm_adj_hess_acc<false>(Hn[st], mu[st], Vx, G);
      tm_adj_dk<false>(J[st], Vx, dk);
      const double wgt = (st == 3) ? 1.0 : 2.0;
      const double cn = (st == 2) ? 1.0 : 0.5;
#pragma unroll
      for (int e = 0; e < NX * NZ; ++e) {
        AB[e] += wgt * dk[e];
        if (st < 3) Vx[e] = ((e / NZ == e % NZ) ? 1.0 : 0.0) + cn * h * dk[e];
      }
    }
#pragma unroll
    for (int e = 0; e < NX * NZ; ++e) AB[e] = ((e / NZ == e % NZ) ? 1.0 : 0.0) + h / 6.0 * AB[e];
    if (order == 2) {
      // Psi <- D' Psi D + G,  D = [A B; 0 I]:   T = Psi D (NZ x NZ), Psi' = D' T (upper triangle)
      double T[NZ * NZ];
#pragma unroll
      for (int r = 0; r < NZ; ++r)
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
          double s = (q >= NX) ? Psi[r <= q ? tm_pair_idx(r, q) : tm_pair_idx(q, r)] : 0.0;
#pragma unroll
          for (int b = 0; b < NX; ++b) s += Psi[r <= b ? tm_pair_idx(r, b) : tm_pair_idx(b, r)] * AB[b * NZ + q];
          T[r * NZ + q] = s;
        }
#pragma unroll
      for (int r = 0; r < NZ; ++r)
#pragma unroll
        for (int q = r; q < NZ; ++q) {
          double s = G[tm_pair_idx(r, q)] + ((r >= NX) ? T[r * NZ + q] : 0.0);
#pragma unroll
          for (int b = 0; b < NX; ++b) s += AB[b * NZ + r] * T[b * NZ + q];
          Psi[tm_pair_idx(r, q)] = s;
        }
      // l <- A' l
      double ln[NX];
#pragma unroll
      for (int b = 0; b < NX; ++b) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) s += AB[a * NZ + b] * l[a];
        ln[b] = s;
      }
#pragma unroll
      for (int b = 0; b < NX; ++b) l[b] = ln[b];
    }
    // R <- R D
    {
      double Rn[NX * NZ];
#pragma unroll
      for (int a = 0; a < NX; ++a)
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
          double s = (q >= NX) ? R[a * NZ + q] : 0.0;
#pragma unroll
          for (int b = 0; b < NX; ++b) s += R[a * NZ + b] * AB[b * NZ + q];
          Rn[a * NZ + q] = s;
        }
#pragma unroll
      for (int e = 0; e < NX * NZ; ++e) R[e] = Rn[e];
    }
  }
#pragma unroll
  for (int e = 0; e < NX * NZ; ++e) rec[NX + e] = R[e];
  if (order == 2) {
#pragma unroll
    for (int e = 0; e < TM_NPAIR; ++e) rec[NX + NX * NZ + e] = Psi[e];
  }
}

#endif  // TMPC_RK4

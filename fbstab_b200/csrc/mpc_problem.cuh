// mpc_problem.cuh -- Problem policy for MPC-structured QPs (one team per
// instance walks the horizon).
//
// Device counterpart of MpcData (reference fbstab/components/mpc_data.cc:17-289,
// mpc_data.h:81-97), RiccatiLinearSolver (riccati_linear_solver.cc:77-344) and
// FullFeasibility (full_feasibility.cc:25-88) for that data class.
//
// The recursion is the reference's, stage by stage and in the same order
// (Rao/Wright/Rawlings Riccati recursion for the barrier-augmented KKT
// system); only the intra-stage work is spread over the team's threads.
// z = [x0;u0;...;xN;uN], l = [l0..lN] (nx each), v,y = nc per stage
// (mpc_data.cc:32-35).  Sign conventions: b = -d, h = -[x0;c0;...;c(N-1)],
// G row-block 0 = [-I 0], row-block i = [A(i-1) B(i-1)] at stage i-1 and -I at
// x(i) (mpc_data.cc:107-151,260-289).
#pragma once

#include "common.cuh"

namespace fbs {

// ---- small dense helpers (column-major, ld = rows), team-cooperative -------

// In-place lower Cholesky (Eigen LLT unblocked order).  false on pivot <= 0.
__device__ inline bool team_chol(const Team& t, double* M, int m) {
  bool ok = true;
  for (int k = 0; k < m; k++) {
    double s = 0.0;
    for (int j = 0; j < k; j++) {
      const double a = M[k + j * m];
      s = fma(a, a, s);
    }
    double x = M[k + k * m] - s;
    if (!(x > 0.0)) ok = false;
    x = sqrt(x);
    for (int i = k + 1 + t.rank(); i < m; i += t.size()) {
      double a = 0.0;
      for (int j = 0; j < k; j++) a = fma(M[i + j * m], M[k + j * m], a);
      M[i + k * m] = (M[i + k * m] - a) / x;
    }
    t.sync();
    if (t.rank() == 0) M[k + k * m] = x;  // after everyone has read the old pivot
  }
  t.sync();
  return ok;
}

// y <- L^-1 x (lower, non-unit, column oriented).  x is destroyed.
__device__ inline void trsv_l(const Team& t, const double* L, int m, double* x,
                              double* y) {
  for (int j = 0; j < m; j++) {
    const double xj = x[j] / L[j + j * m];
    if (t.rank() == 0) y[j] = xj;
    for (int i = j + 1 + t.rank(); i < m; i += t.size())
      x[i] = fma(-L[i + j * m], xj, x[i]);
    t.sync();
  }
}
// y <- L^-T x.  x is destroyed.
__device__ inline void trsv_lt(const Team& t, const double* L, int m, double* x,
                               double* y) {
  for (int i = m - 1; i >= 0; i--) {
    const double xi = x[i] / L[i + i * m];
    if (t.rank() == 0) y[i] = xi;
    for (int r = t.rank(); r < i; r += t.size())
      x[r] = fma(-L[i + r * m], xi, x[r]);
    t.sync();
  }
}
// one thread: row r of X(rows x m) <- (src row) * L^-T  (forward substitution)
__device__ __forceinline__ void row_trsm_lt(const double* L, int m,
                                            const double* src, double* X,
                                            int rows, int r) {
  for (int j = 0; j < m; j++) {
    double s = src[r + j * rows];
    for (int k = 0; k < j; k++) s = fma(-X[r + k * rows], L[j + k * m], s);
    X[r + j * rows] = s / L[j + j * m];
  }
}

struct MpcProblem {
  int N, nx, nu, nc, nz, nl, nv;
  // this instance's sequences
  const double *Q, *R, *S, *q, *r, *A, *B, *c, *E, *L, *d, *x0;
  // workspace (per CTA)
  double *gamma, *mus, *Gam, *tv;      // nv each
  double *Ls, *Ms, *AMs, *SMs, *SGs, *Ps;  // per-stage factors, N+1 each
  double *Qt, *Rt, *St, *Linv;         // stage temporaries
  double *r1, *r2, *hs, *ths, *txs, *tus;  // nz, nl, nl, nl, nl, (N+1)*nu
  double *sa, *sb, *sc;                // max(nx,nu) each
  double* tzs;                         // nz scratch

  __device__ __forceinline__ int ns() const { return nx + nu; }
  __device__ __forceinline__ double b(int i) const { return -d[i]; }
  __device__ __forceinline__ const double* Qi(int i) const { return Q + (size_t)i * nx * nx; }
  __device__ __forceinline__ const double* Ri(int i) const { return R + (size_t)i * nu * nu; }
  __device__ __forceinline__ const double* Si(int i) const { return S + (size_t)i * nu * nx; }
  __device__ __forceinline__ const double* Ai(int i) const { return A + (size_t)i * nx * nx; }
  __device__ __forceinline__ const double* Bi(int i) const { return B + (size_t)i * nx * nu; }
  __device__ __forceinline__ const double* Ei(int i) const { return E + (size_t)i * nc * nx; }
  __device__ __forceinline__ const double* Lci(int i) const { return L + (size_t)i * nc * nu; }

  // mpc_data.h:89-97
  __device__ double forcing_norm(const Team& t) const {
    double s[1] = {0.0};
    for (int i = t.rank(); i < (N + 1) * nx; i += t.size()) s[0] += q[i] * q[i];
    for (int i = t.rank(); i < (N + 1) * nu; i += t.size()) s[0] += r[i] * r[i];
    for (int i = t.rank(); i < (N + 1) * nc; i += t.size()) s[0] += d[i] * d[i];
    for (int i = t.rank(); i < nx; i += t.size()) s[0] += x0[i] * x0[i];
    for (int i = t.rank(); i < N * nx; i += t.size()) s[0] += c[i] * c[i];
    team_sum(t, s);
    return sqrt(s[0]);
  }

  // (E(i) x(i) + L(i) u(i))[k]   -- one entry of A_qp z, mpc_data.cc:66-105
  __device__ __forceinline__ double Az_entry(const double* z, int i, int k) const {
    const double* Em = Ei(i);
    const double* Lm = Lci(i);
    const double* xi = z + (size_t)i * ns();
    const double* ui = xi + nx;
    double s = 0.0;
    for (int cc = 0; cc < nx; cc++) s = fma(Em[k + cc * nc], xi[cc], s);
    double s2 = 0.0;
    for (int cc = 0; cc < nu; cc++) s2 = fma(Lm[k + cc * nc], ui[cc], s2);
    return s + s2;
  }

  // y = b - A z
  __device__ void margin(const Team& t, const double* z, double* y) const {
    for (int e = t.rank(); e < nv; e += t.size()) {
      const int i = e / nc, k = e - i * nc;
      y[e] = -d[e] - Az_entry(z, i, k);
    }
    t.sync();
  }

  // (H z)[idx], mpc_data.cc:17-64
  __device__ __forceinline__ double Hz_entry(const double* z, int i, int rr) const {
    const double* xi = z + (size_t)i * ns();
    const double* ui = xi + nx;
    double s1 = 0.0, s2 = 0.0;
    if (rr < nx) {
      const double* Qm = Qi(i);
      const double* Sm = Si(i);
      for (int cc = 0; cc < nx; cc++) s1 = fma(Qm[rr + cc * nx], xi[cc], s1);
      for (int cc = 0; cc < nu; cc++) s2 = fma(Sm[cc + rr * nu], ui[cc], s2);
    } else {
      const int ru = rr - nx;
      const double* Sm = Si(i);
      const double* Rm = Ri(i);
      for (int cc = 0; cc < nx; cc++) s1 = fma(Sm[ru + cc * nu], xi[cc], s1);
      for (int cc = 0; cc < nu; cc++) s2 = fma(Rm[ru + cc * nu], ui[cc], s2);
    }
    return s1 + s2;
  }
  // acc + (G' l)[idx] in the reference's order: first the -l(i) term, then
  // the A(i)'/B(i)' l(i+1) product (mpc_data.cc:153-199)
  __device__ __forceinline__ double add_GTl(double acc, const double* l, int i,
                                            int rr) const {
    if (rr < nx) {
      acc += -l[(size_t)i * nx + rr];
      if (i < N) {
        const double* Am = Ai(i);
        const double* lp = l + (size_t)(i + 1) * nx;
        double s = 0.0;
        for (int cc = 0; cc < nx; cc++) s = fma(Am[cc + rr * nx], lp[cc], s);
        acc += s;
      }
      return acc;
    }
    if (i < N) {
      const int ru = rr - nx;
      const double* Bm = Bi(i);
      const double* lp = l + (size_t)(i + 1) * nx;
      double s = 0.0;
      for (int cc = 0; cc < nx; cc++) s = fma(Bm[cc + ru * nx], lp[cc], s);
      acc += s;
    }
    return acc;
  }
  // (A' v)[idx], mpc_data.cc:201-240
  __device__ __forceinline__ double ATv_entry(const double* v, int i, int rr) const {
    const double* vi = v + (size_t)i * nc;
    double s = 0.0;
    if (rr < nx) {
      const double* Em = Ei(i) + (size_t)rr * nc;
      for (int k = 0; k < nc; k++) s = fma(Em[k], vi[k], s);
    } else {
      const double* Lm = Lci(i) + (size_t)(rr - nx) * nc;
      for (int k = 0; k < nc; k++) s = fma(Lm[k], vi[k], s);
    }
    return s;
  }
  // (A(i-1) x(i-1) + B(i-1) u(i-1))[rr], i >= 1   (mpc_data.cc:123-140)
  __device__ __forceinline__ double AB_entry(const double* z, int i, int rr) const {
    const double* xm = z + (size_t)(i - 1) * ns();
    const double* um = xm + nx;
    const double* Am = Ai(i - 1);
    const double* Bm = Bi(i - 1);
    double s1 = 0.0, s2 = 0.0;
    for (int cc = 0; cc < nx; cc++) s1 = fma(Am[rr + cc * nx], xm[cc], s1);
    for (int cc = 0; cc < nu; cc++) s2 = fma(Bm[rr + cc * nx], um[cc], s2);
    return s1 + s2;
  }
  // (G z)[i*nx + rr], mpc_data.cc:107-151
  __device__ __forceinline__ double Gz_entry(const double* z, int i, int rr) const {
    if (i == 0) return -z[rr];
    return AB_entry(z, i, rr) - z[(size_t)i * ns() + rr];
  }
  __device__ __forceinline__ double f_entry(int i, int rr) const {
    return rr < nx ? q[(size_t)i * nx + rr] : r[(size_t)i * nu + rr - nx];
  }
  __device__ __forceinline__ double h_entry(int i, int rr) const {
    return i == 0 ? -x0[rr] : -c[(size_t)(i - 1) * nx + rr];
  }

  // tz = ((f + Hz) + G'l) + A'v ; tl = h - Gz
  __device__ void kkt(const Team& t, const Vars& x, double* oz, double* ol) const {
    const int nsv = ns();
    for (int e = t.rank(); e < nz + nl; e += t.size()) {
      if (e < nz) {
        const int i = e / nsv, rr = e - i * nsv;
        double v = f_entry(i, rr) + Hz_entry(x.z, i, rr);
        v = add_GTl(v, x.l, i, rr);
        v += ATv_entry(x.v, i, rr);
        oz[e] = v;
      } else {
        const int k = e - nz;
        const int i = k / nx, rr = k - i * nx;
        // l = h ; l += -(A x + B u) ; l += x(i)   (gemvG with a = -1)
        if (i == 0)
          ol[k] = h_entry(0, rr) + x.z[rr];
        else
          ol[k] = (h_entry(i, rr) - AB_entry(x.z, i, rr)) +
                  x.z[(size_t)i * nsv + rr];
      }
    }
    t.sync();
  }

  // RiccatiLinearSolver::Initialize, riccati_linear_solver.cc:77-210
  __device__ bool factor(const Team& t, const Vars& x, const Vars& xbar,
                         double sigma, double alpha) {
    for (int i = t.rank(); i < nv; i += t.size()) {
      const double ys = x.y[i] + sigma * (x.v[i] - xbar.v[i]);
      double ga, mu;
      pfb_barrier(ys, x.v[i], alpha, sigma, &ga, &mu);
      gamma[i] = ga;
      mus[i] = mu;
      Gam[i] = ga / mu;
    }
    const int nxx = nx * nx, nuu = nu * nu, nux = nu * nx;
    // L(0) = sqrt(sigma) I, :127
    const double rs = sqrt(sigma);
    for (int e = t.rank(); e < nxx; e += t.size())
      Ls[e] = (e % nx == e / nx) ? rs : 0.0;
    t.sync();
    bool ok = true;
    for (int i = 0; i <= N; i++) {
      double* Li = Ls + (size_t)i * nxx;
      double* Mi = Ms + (size_t)i * nxx;
      double* AMi = AMs + (size_t)i * nxx;
      double* SMi = SMs + (size_t)i * nux;
      double* SGi = SGs + (size_t)i * nuu;
      double* Pi = Ps + (size_t)i * nux;
      const double* Em = Ei(i);
      const double* Lm = Lci(i);
      const double* Gi = Gam + (size_t)i * nc;
      // barrier-augmented stage Hessian (:102-123) and Linv = inv(L L') (:142-144)
      const int work = nxx + nuu + nux + nx;
      for (int e = t.rank(); e < work; e += t.size()) {
        if (e < nxx) {
          const int rr = e % nx, cc = e / nx;
          if (rr >= cc) {
            double s = 0.0;
            for (int k = 0; k < nc; k++)
              s = fma(Em[k + rr * nc], Gi[k] * Em[k + cc * nc], s);
            Qt[e] = (Qi(i)[e] + (rr == cc ? sigma : 0.0)) + s;
          }
        } else if (e < nxx + nuu) {
          const int f = e - nxx;
          const int rr = f % nu, cc = f / nu;
          if (rr >= cc) {
            double s = 0.0;
            for (int k = 0; k < nc; k++)
              s = fma(Lm[k + rr * nc], Gi[k] * Lm[k + cc * nc], s);
            Rt[f] = (Ri(i)[f] + (rr == cc ? sigma : 0.0)) + s;
          }
        } else if (e < nxx + nuu + nux) {
          const int f = e - nxx - nuu;
          const int rr = f % nu, cc = f / nu;
          double s = 0.0;
          for (int k = 0; k < nc; k++)
            s = fma(Lm[k + rr * nc], Gi[k] * Em[k + cc * nc], s);
          St[f] = Si(i)[f] + s;
        } else {
          // column cc of inv(L L'): forward then backward substitution
          const int cc = e - nxx - nuu - nux;
          double* w = Linv + (size_t)cc * nx;
          for (int k = 0; k < nx; k++) w[k] = (k == cc) ? 1.0 : 0.0;
          for (int j = 0; j < nx; j++) {
            w[j] /= Li[j + j * nx];
            const double wj = w[j];
            for (int k = j + 1; k < nx; k++) w[k] = fma(-Li[k + j * nx], wj, w[k]);
          }
          for (int k = nx - 1; k >= 0; k--) {
            double s = w[k];
            for (int j = k + 1; j < nx; j++) s = fma(-Li[j + k * nx], w[j], s);
            w[k] = s / Li[k + k * nx];
          }
        }
      }
      t.sync();
      // M = chol(Qt + Linv), :145-147
      for (int e = t.rank(); e < nxx; e += t.size()) {
        const int rr = e % nx, cc = e / nx;
        if (rr >= cc) Mi[e] = Qt[e] + Linv[e];
      }
      t.sync();
      ok = team_chol(t, Mi, nx) && ok;
      // AM = A M^-T, SM = St M^-T, :149-161
      for (int w = t.rank(); w < nx + nu; w += t.size()) {
        if (w < nx) {
          if (i < N) row_trsm_lt(Mi, nx, Ai(i), AMi, nx, w);
        } else {
          row_trsm_lt(Mi, nx, St, SMi, nu, w - nx);
        }
      }
      t.sync();
      // SG = chol(Rt - SM SM'), :163-166
      for (int e = t.rank(); e < nuu; e += t.size()) {
        const int rr = e % nu, cc = e / nu;
        if (rr >= cc) {
          double s = 0.0;
          for (int k = 0; k < nx; k++) s = fma(SMi[rr + k * nu], SMi[cc + k * nu], s);
          SGi[e] = Rt[e] - s;
        }
      }
      t.sync();
      ok = team_chol(t, SGi, nu) && ok;
      if (i == N) break;
      // P = (AM SM' - B) SG^-T, :170-175
      for (int rr = t.rank(); rr < nx; rr += t.size()) {
        const double* Bm = Bi(i);
        for (int j = 0; j < nu; j++) {
          double s = 0.0;
          for (int k = 0; k < nx; k++) s = fma(AMi[rr + k * nx], SMi[j + k * nu], s);
          Pi[rr + j * nx] = s - Bm[rr + j * nx];
        }
        row_trsm_lt(SGi, nu, Pi, Pi, nx, rr);
      }
      t.sync();
      // L(i+1) = chol(sigma I + P P' + AM AM'), :179-183
      double* Ln = Li + nxx;
      for (int e = t.rank(); e < nxx; e += t.size()) {
        const int rr = e % nx, cc = e / nx;
        if (rr >= cc) {
          double s1 = 0.0, s2 = 0.0;
          for (int k = 0; k < nu; k++) s1 = fma(Pi[rr + k * nx], Pi[cc + k * nx], s1);
          for (int k = 0; k < nx; k++) s2 = fma(AMi[rr + k * nx], AMi[cc + k * nx], s2);
          Ln[e] = ((rr == cc ? sigma : 0.0) + s1) + s2;
        } else {
          Ln[e] = 0.0;
        }
      }
      t.sync();
      ok = team_chol(t, Ln, nx) && ok;
    }
    return ok;
  }

  // y(rows) = M(rows x cols) x, one entry per thread; no sync
  __device__ __forceinline__ double mv_row(const double* M, int rows, int cols,
                                           const double* xx, int rr) const {
    double s = 0.0;
    for (int cc = 0; cc < cols; cc++) s = fma(M[rr + cc * rows], xx[cc], s);
    return s;
  }
  // (M' x)[cc]
  __device__ __forceinline__ double mtv_col(const double* M, int rows,
                                            const double* xx, int cc) const {
    double s = 0.0;
    for (int rr = 0; rr < rows; rr++) s = fma(M[rr + cc * rows], xx[rr], s);
    return s;
  }

  // RiccatiLinearSolver::Solve on r = -(rz,rl,rv), riccati_linear_solver.cc:212-344
  __device__ void solve(const Team& t, const double* rz, const double* rl,
                        const double* rv, const Vars& dx) {
    const int nsv = ns();
    const int nxx = nx * nx, nuu = nu * nu, nux = nu * nx;
    // r3 = r.v ./ mus ; r1 = r.z - A' r3 ; r2 = -r.l   (:222-225)
    for (int i = t.rank(); i < nv; i += t.size()) tv[i] = (-rv[i]) / mus[i];
    for (int i = t.rank(); i < nl; i += t.size()) r2[i] = rl[i];
    t.sync();
    for (int e = t.rank(); e < nz; e += t.size()) {
      const int i = e / nsv, rr = e - i * nsv;
      r1[e] = (-rz[e]) - ATv_entry(tv, i, rr);
    }
    t.sync();
    // base case :232-236
    for (int k = t.rank(); k < nx; k += t.size()) {
      ths[k] = r2[k];
      sa[k] = r2[k];
    }
    t.sync();
    trsv_l(t, Ls, nx, sa, sb);
    trsv_lt(t, Ls, nx, sb, sc);
    for (int k = t.rank(); k < nx; k += t.size()) hs[k] = sc[k] - r1[k];
    t.sync();
    // forward recursion :239-262
    for (int i = 0; i < N; i++) {
      const double* Mi = Ms + (size_t)i * nxx;
      const double* AMi = AMs + (size_t)i * nxx;
      const double* SMi = SMs + (size_t)i * nux;
      const double* SGi = SGs + (size_t)i * nuu;
      const double* Pi = Ps + (size_t)i * nux;
      const double* Ln = Ls + (size_t)(i + 1) * nxx;
      double* tx = txs + (size_t)i * nx;
      double* tu = tus + (size_t)i * nu;
      for (int k = t.rank(); k < nx; k += t.size()) sa[k] = hs[(size_t)i * nx + k];
      t.sync();
      trsv_l(t, Mi, nx, sa, tx);  // tx = M^-1 h
      for (int k = t.rank(); k < nu; k += t.size())
        sb[k] = mv_row(SMi, nu, nx, tx, k) + r1[(size_t)i * nsv + nx + k];
      t.sync();
      trsv_l(t, SGi, nu, sb, tu);  // tu = SG^-1 (SM tx + ru)
      for (int k = t.rank(); k < nx; k += t.size()) {
        const double v = (mv_row(Pi, nx, nu, tu, k) + mv_row(AMi, nx, nx, tx, k)) +
                         r2[(size_t)(i + 1) * nx + k];
        ths[(size_t)(i + 1) * nx + k] = v;
        sa[k] = v;
      }
      t.sync();
      trsv_l(t, Ln, nx, sa, sb);
      trsv_lt(t, Ln, nx, sb, sc);
      for (int k = t.rank(); k < nx; k += t.size())
        hs[(size_t)(i + 1) * nx + k] = sc[k] - r1[(size_t)(i + 1) * nsv + k];
      t.sync();
    }
    // terminal stage :267-285
    {
      const double* Mi = Ms + (size_t)N * nxx;
      const double* SMi = SMs + (size_t)N * nux;
      const double* SGi = SGs + (size_t)N * nuu;
      const double* LN = Ls + (size_t)N * nxx;
      double* tx = txs + (size_t)N * nx;
      double* uN = dx.z + (size_t)N * nsv + nx;
      double* xN = dx.z + (size_t)N * nsv;
      double* lN = dx.l + (size_t)N * nx;
      for (int k = t.rank(); k < nx; k += t.size()) sa[k] = hs[(size_t)N * nx + k];
      t.sync();
      trsv_l(t, Mi, nx, sa, tx);
      for (int k = t.rank(); k < nu; k += t.size())
        sb[k] = mv_row(SMi, nu, nx, tx, k) + r1[(size_t)N * nsv + nx + k];
      t.sync();
      trsv_l(t, SGi, nu, sb, sc);
      trsv_lt(t, SGi, nu, sc, uN);
      for (int k = t.rank(); k < nx; k += t.size())
        sa[k] = tx[k] + mtv_col(SMi, nu, uN, k);
      t.sync();
      trsv_lt(t, Mi, nx, sa, sb);
      for (int k = t.rank(); k < nx; k += t.size()) {
        const double xv = -sb[k];
        xN[k] = xv;
        sa[k] = xv + ths[(size_t)N * nx + k];
      }
      t.sync();
      trsv_l(t, LN, nx, sa, sb);
      trsv_lt(t, LN, nx, sb, sc);
      for (int k = t.rank(); k < nx; k += t.size()) lN[k] = -sc[k];
      t.sync();
    }
    // backward recursion :297-327 (tx and SG^-1(SM tx + ru) were kept from the
    // forward pass; the reference recomputes the same values)
    for (int i = N - 1; i >= 0; i--) {
      const double* Mi = Ms + (size_t)i * nxx;
      const double* AMi = AMs + (size_t)i * nxx;
      const double* SMi = SMs + (size_t)i * nux;
      const double* SGi = SGs + (size_t)i * nuu;
      const double* Pi = Ps + (size_t)i * nux;
      const double* Li = Ls + (size_t)i * nxx;
      const double* tx = txs + (size_t)i * nx;
      const double* tu = tus + (size_t)i * nu;
      const double* lp = dx.l + (size_t)(i + 1) * nx;
      double* ui = dx.z + (size_t)i * nsv + nx;
      double* xi = dx.z + (size_t)i * nsv;
      double* li = dx.l + (size_t)i * nx;
      for (int k = t.rank(); k < nu; k += t.size())
        sa[k] = tu[k] + mtv_col(Pi, nx, lp, k);
      t.sync();
      trsv_lt(t, SGi, nu, sa, ui);
      for (int k = t.rank(); k < nx; k += t.size())
        sa[k] = (tx[k] + mtv_col(SMi, nu, ui, k)) + mtv_col(AMi, nx, lp, k);
      t.sync();
      trsv_lt(t, Mi, nx, sa, sb);
      for (int k = t.rank(); k < nx; k += t.size()) {
        const double xv = -sb[k];
        xi[k] = xv;
        sa[k] = ths[(size_t)i * nx + k] + xv;
      }
      t.sync();
      trsv_l(t, Li, nx, sa, sb);
      trsv_lt(t, Li, nx, sb, sc);
      for (int k = t.rank(); k < nx; k += t.size()) li[k] = -sc[k];
      t.sync();
    }
    // dv = (rv + gamma .* A dz) ./ mus ; dy = b - A dz   (:331-341)
    for (int e = t.rank(); e < nv; e += t.size()) {
      const int i = e / nc, k = e - i * nc;
      const double s = Az_entry(dx.z, i, k);
      dx.v[e] = ((-rv[e]) + gamma[e] * s) / mus[e];
      dx.y[e] = (-s) + (-d[e]);
    }
    t.sync();
  }

  // FullFeasibility::CheckFeasibility, full_feasibility.cc:25-88
  __device__ int feasibility(const Team& t, const Vars& dx, double tol) {
    const int nsv = ns();
    double mx[4] = {-INFINITY, 0.0, 0.0, 0.0};
    double mp[3] = {0.0, 0.0, 0.0};
    double sm[2] = {0.0, 0.0};
    for (int e = t.rank(); e < nv; e += t.size()) {
      const int i = e / nc, k = e - i * nc;
      mx[0] = fmax(mx[0], Az_entry(dx.z, i, k));
      mp[1] = fmax(mp[1], fabs(dx.v[e]));
      sm[1] += (-d[e]) * dx.v[e];
    }
    for (int e = t.rank(); e < nl; e += t.size()) {
      const int i = e / nx, rr = e - i * nx;
      mx[1] = fmax(mx[1], fabs(Gz_entry(dx.z, i, rr)));
      mp[2] = fmax(mp[2], fabs(dx.l[e]));
      sm[1] += h_entry(i, rr) * dx.l[e];
    }
    for (int e = t.rank(); e < nz; e += t.size()) {
      const int i = e / nsv, rr = e - i * nsv;
      mx[2] = fmax(mx[2], fabs(Hz_entry(dx.z, i, rr)));
      mx[3] = fmax(mx[3], fabs(dx.z[e]));
      sm[0] += f_entry(i, rr) * dx.z[e];
      const double p = add_GTl(ATv_entry(dx.v, i, rr), dx.l, i, rr);
      mp[0] = fmax(mp[0], fabs(p));
    }
    team_max(t, mx);
    team_max(t, mp);
    team_sum(t, sm);
    t.sync();
    const double w = mx[3];
    const bool dual_infeasible = (mx[0] <= w * tol) && (mx[1] <= tol * w) &&
                                 (mx[2] <= tol * w) && (sm[0] < 0.0) &&
                                 (w > 1e-14);
    const double u = fmax(mp[1], mp[2]);
    const bool primal_infeasible = (mp[0] <= tol * u) && (sm[1] < 0.0);
    return (primal_infeasible ? 1 : 0) + (dual_infeasible ? 2 : 0);
  }
};

}  // namespace fbs

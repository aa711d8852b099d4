// dense_problem.cuh -- Problem policy for dense QPs of arbitrary size
// (one CTA per instance, matrices streamed from global memory / L2).
//
// Device counterpart of DenseData (reference fbstab/components/dense_data.cc:12-41),
// DenseCholeskySolver (dense_cholesky_solver.cc:32-148) and FullFeasibility
// (full_feasibility.cc:25-88) for that data class.
//
// Linear solver.  The reference factors K = [E G'; G -sigma I],
// E = H + sigma I + A' Gamma A, with Eigen's diagonally pivoted LDL'.  That
// pivot rule looks only at the ORIGINAL diagonal, so it always eliminates the
// E block (diagonal >= sigma + H_ii > sigma) before the -sigma I block: it is
// a block elimination of E followed by the Schur complement.  The engine does
// the same elimination without the intra-block permutation (unpivoted LDL' of
// the quasi-definite K, which exists for every symmetric permutation).
#pragma once

#include "common.cuh"

namespace fbs {

struct DenseProblem {
  int nz, nl, nv, n;  // n = nz + nl
  const double *H, *f, *G, *h, *A, *bvec;  // this instance, column-major
  // workspace (global, per CTA)
  double* K;      // n*n, lower triangle used
  double* r1;     // n
  double* r2;     // nv   (Gamma during factor, scratch during solve)
  double* gamma;  // nv
  double* mus;    // nv
  double* tmp;    // n
  double* tz;     // nz  scratch (feasibility)

  __device__ __forceinline__ double b(int i) const { return bvec[i]; }
  __device__ __forceinline__ double fvec(int i) const { return f[i]; }
  __device__ __forceinline__ double hvec(int i) const { return h[i]; }

  // dense_data.h:72-73
  __device__ double forcing_norm(const Team& t) const {
    double s[1] = {0.0};
    for (int i = t.rank(); i < nv; i += t.size()) s[0] += bvec[i] * bvec[i];
    for (int i = t.rank(); i < nz; i += t.size()) s[0] += f[i] * f[i];
    for (int i = t.rank(); i < nl; i += t.size()) s[0] += h[i] * h[i];
    team_sum(t, s);
    return sqrt(s[0]);
  }

  // y = b - A z  (full_variable.cc:47-53)
  __device__ void margin(const Team& t, const double* z, double* y) const {
    for (int i = t.rank(); i < nv; i += t.size()) {
      double s = 0.0;
      for (int j = 0; j < nz; j++) s = fma(A[i + (size_t)j * nv], z[j], s);
      y[i] = bvec[i] - s;
    }
    t.sync();
  }

  // tz = ((f + Hz) + G'l) + A'v ; tl = h - Gz   (full_residual.cc:52-63)
  __device__ void kkt(const Team& t, const Vars& x, double* oz, double* ol) const {
    FBS_LAP(15);
    for (int i = t.rank(); i < n; i += t.size()) {
      if (i < nz) {
        double s = 0.0;
        for (int j = 0; j < nz; j++) s = fma(H[i + (size_t)j * nz], x.z[j], s);
        oz[i] = f[i] + s;
      } else {
        const int k = i - nz;
        double s = 0.0;
        for (int j = 0; j < nz; j++) s = fma(G[k + (size_t)j * nl], x.z[j], s);
        ol[k] = h[k] - s;
      }
    }
    t.sync();
    FBS_LAP(11);
    for (int i = t.warp(); i < nz; i += t.nwarps()) {
      double s1 = 0.0, s2 = 0.0;
      const double* g = G + (size_t)i * nl;
      const double* a = A + (size_t)i * nv;
      for (int k = t.lane(); k < nl; k += 32) s1 = fma(g[k], x.l[k], s1);
      for (int k = t.lane(); k < nv; k += 32) s2 = fma(a[k], x.v[k], s2);
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (t.lane() == 0) oz[i] = (oz[i] + s1) + s2;
    }
    t.sync();
    FBS_LAP(12);
  }

  // LinearSolver::Initialize, dense_cholesky_solver.cc:32-79
  __device__ bool factor(const Team& t, const Vars& x, const Vars& xbar,
                         double sigma, double alpha) {
    double* Gam = r2;
    for (int i = t.rank(); i < nv; i += t.size()) {
      const double ys = x.y[i] + sigma * (x.v[i] - xbar.v[i]);
      double ga, mu;
      pfb_barrier(ys, x.v[i], alpha, sigma, &ga, &mu);
      gamma[i] = ga;
      mus[i] = mu;
      Gam[i] = div_nr(ga, mu);
    }
    t.sync();
    // E = (H + sigma I) + A' (Gamma A), lower triangle -> K(0:nz,0:nz)
    for (int e = t.rank(); e < nz * nz; e += t.size()) {
      const int i = e % nz, j = e / nz;
      if (i < j) continue;
      const double* ai = A + (size_t)i * nv;
      const double* aj = A + (size_t)j * nv;
      double s = 0.0;
      for (int k = 0; k < nv; k++) s = fma(ai[k], Gam[k] * aj[k], s);
      K[i + (size_t)j * n] = (H[i + (size_t)j * nz] + (i == j ? sigma : 0.0)) + s;
    }
    // [G  -sigma I] rows
    for (int e = t.rank(); e < nl * n; e += t.size()) {
      const int r = e % nl, c = e / nl;
      double val;
      if (c < nz)
        val = G[r + (size_t)c * nl];
      else
        val = (c - nz == r) ? -sigma : 0.0;
      K[nz + r + (size_t)c * n] = val;
    }
    t.sync();
    // left-looking LDL' (Eigen LDLT.h unblocked, without the permutation)
    bool ok = true;
    for (int kk = 0; kk < n; kk++) {
      for (int j = t.rank(); j < kk; j += t.size())
        tmp[j] = K[j + (size_t)j * n] * K[kk + (size_t)j * n];
      t.sync();
      for (int i = kk + t.rank(); i < n; i += t.size()) {
        double s = 0.0;
        for (int j = 0; j < kk; j++) s = fma(K[i + (size_t)j * n], tmp[j], s);
        K[i + (size_t)kk * n] -= s;
      }
      t.sync();
      const double d = K[kk + (size_t)kk * n];
      if (!(fabs(d) > 0.0)) ok = false;  // zero / NaN pivot
      for (int i = kk + 1 + t.rank(); i < n; i += t.size())
        K[i + (size_t)kk * n] /= d;
      t.sync();
    }
    return ok;
  }

  // LinearSolver::Solve with r = -(rz,rl,rv), dense_cholesky_solver.cc:81-127
  __device__ void solve(const Team& t, const double* rz, const double* rl,
                        const double* rv, const Vars& dx) {
    for (int i = t.rank(); i < nv; i += t.size()) r2[i] = div_nr(-rv[i], mus[i]);
    for (int i = t.rank(); i < nl; i += t.size()) r1[nz + i] = -(-rl[i]);
    t.sync();
    for (int i = t.warp(); i < nz; i += t.nwarps()) {
      const double* a = A + (size_t)i * nv;
      double s = 0.0;
      for (int k = t.lane(); k < nv; k += 32) s = fma(a[k], r2[k], s);
      s = warp_sum(s);
      if (t.lane() == 0) r1[i] = (-rz[i]) - s;
    }
    t.sync();
    // L^-1 (unit lower, column oriented)
    for (int j = 0; j < n; j++) {
      const double xj = r1[j];
      for (int i = j + 1 + t.rank(); i < n; i += t.size())
        r1[i] -= K[i + (size_t)j * n] * xj;
      t.sync();
    }
    // D^-1 (pseudo-inverse threshold as Eigen's LDLT::solve)
    for (int i = t.rank(); i < n; i += t.size()) {
      const double d = K[i + (size_t)i * n];
      r1[i] = (fabs(d) > 2.2250738585072014e-308) ? r1[i] / d : 0.0;
    }
    t.sync();
    // L^-T (row oriented on the stored lower factor)
    for (int j = n - 1; j > 0; j--) {
      const double xj = r1[j];
      for (int i = t.rank(); i < j; i += t.size())
        r1[i] -= K[j + (size_t)i * n] * xj;
      t.sync();
    }
    for (int i = t.rank(); i < nz; i += t.size()) dx.z[i] = r1[i];
    for (int i = t.rank(); i < nl; i += t.size()) dx.l[i] = r1[nz + i];
    t.sync();
    // dv = (rv + gamma .* (A dz)) ./ mus ; dy = b - A dz
    for (int i = t.rank(); i < nv; i += t.size()) {
      double s = 0.0;
      for (int j = 0; j < nz; j++) s = fma(A[i + (size_t)j * nv], dx.z[j], s);
      dx.v[i] = div_nr(gamma[i] * s + (-rv[i]), mus[i]);
      dx.y[i] = bvec[i] - s;
    }
    t.sync();
  }

  // FullFeasibility::CheckFeasibility, full_feasibility.cc:25-88.
  // 0 feasible, 1 primal infeasible, 2 dual infeasible, 3 both.
  __device__ int feasibility(const Team& t, const Vars& dx, double tol) {
    // d1 = max(A dz), d2 = |G dz|inf, d3 = |H dz|inf, d4 = f'dz, w = |dz|inf
    double mx[4] = {-INFINITY, 0.0, 0.0, 0.0};
    double sm[2] = {0.0, 0.0};
    for (int i = t.rank(); i < nv; i += t.size()) {
      double s = 0.0;
      for (int j = 0; j < nz; j++) s = fma(A[i + (size_t)j * nv], dx.z[j], s);
      mx[0] = fmax(mx[0], s);
    }
    for (int i = t.rank(); i < nl; i += t.size()) {
      double s = 0.0;
      for (int j = 0; j < nz; j++) s = fma(G[i + (size_t)j * nl], dx.z[j], s);
      mx[1] = fmax(mx[1], fabs(s));
    }
    for (int i = t.rank(); i < nz; i += t.size()) {
      double s = 0.0;
      for (int j = 0; j < nz; j++) s = fma(H[i + (size_t)j * nz], dx.z[j], s);
      mx[2] = fmax(mx[2], fabs(s));
      mx[3] = fmax(mx[3], fabs(dx.z[i]));
      sm[0] += f[i] * dx.z[i];
    }
    // p1 = |A'dv + G'dl|inf
    for (int i = t.warp(); i < nz; i += t.nwarps()) {
      double s1 = 0.0, s2 = 0.0;
      const double* g = G + (size_t)i * nl;
      const double* a = A + (size_t)i * nv;
      for (int k = t.lane(); k < nv; k += 32) s1 = fma(a[k], dx.v[k], s1);
      for (int k = t.lane(); k < nl; k += 32) s2 = fma(g[k], dx.l[k], s2);
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (t.lane() == 0) tz[i] = s1 + s2;
    }
    t.sync();
    double mp[3] = {0.0, 0.0, 0.0};  // p1, |dv|inf, |dl|inf
    for (int i = t.rank(); i < nz; i += t.size()) mp[0] = fmax(mp[0], fabs(tz[i]));
    for (int i = t.rank(); i < nv; i += t.size()) {
      mp[1] = fmax(mp[1], fabs(dx.v[i]));
      sm[1] += bvec[i] * dx.v[i];
    }
    for (int i = t.rank(); i < nl; i += t.size()) {
      mp[2] = fmax(mp[2], fabs(dx.l[i]));
      sm[1] += h[i] * dx.l[i];
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

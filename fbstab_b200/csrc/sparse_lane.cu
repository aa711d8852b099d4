// sparse_lane.cu -- lane-per-instance FBstab for batches of SPARSE QPs that share
// one sparsity pattern (FBstabSparse).
//
// The reference plans "general sparse matrix components" (ROADMAP.md:10) on an
// LDL' of the quasi-definite Newton matrix behind the interface its
// tools/qdldl/qdldl_wrapper.h:19-84 sketches (analysis once, Factor, Solve); the
// algorithm around them is FBstabAlgorithm unchanged
// (fbstab_algorithm-impl.h:113-304; data concept abstract_components.h:24-62).
//
// B200 design.  With a common pattern the control flow of EVERYTHING -- the
// gather mat-vecs, the assembly of K, the up-looking numeric LDL' and the
// triangular solves -- depends on the pattern only.  So every LANE owns one
// instance and the 32 lanes of a warp walk the same integer tables
// (sparse_symbolic.h, built once on the host: broadcast loads) while their values
// live lane-interleaved in a per-warp global workspace (element e of lane j at
// ws[e*32 + j]: every access of the warp is one coalesced 256-byte transaction,
// no atomics, no divergence inside a sweep).  The FBstab state machine is the
// per-lane phase machine of lane_engine.cuh, shared with mpc_lane.cu: lanes pull
// instances from the global counter on their own, converged instances drop out.
//
// Newton system (order [z; l; w], w = Gamma^1/2 A dz, so that an inactive
// constraint, gamma = 0, is a harmless -1 pivot instead of a division by zero):
//   [ H + sigma I   G'       (Gamma^1/2 A)' ] [dz]   [ rz - A'(rv / mu) ]
//   [ G            -sigma I   0             ] [dl] = [ -rl              ]
//   [ Gamma^1/2 A   0        -I             ] [w ]   [ 0                ]
//   dv = (rv + gamma .* (A dz)) ./ mu,  dy = b - A dz
// (r = minus the residual; the same reduction as dense_cholesky_solver.cc:81-127
// with the A' Gamma A block left unformed to keep K sparse).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.cuh"
#include "lane_engine.cuh"
#include "sparse_lane.h"

namespace fbs {
namespace {

struct SparseArgs {
  SparseDev d;
  int batch;
  const double *Hx, *f, *Gx, *h, *Ax, *b;
  double *z, *l, *v, *y;
  fbstab_out* out;
  double* ws;        // per-warp workspace base
  size_t ws_stride;  // doubles per warp
  int* counter;
  fbstab_options opts;
  int warps;
  int lanes;         // instances per warp (32, 16 or 8: see SparseLaneLanes)
};

struct SparseLane {
  // ids of the iterate blocks (lane_engine.cuh)
  static constexpr int O_XK = 0, O_XI = 1, O_DX = 2;
  SparseDev d;
  int nz, nl, nv, n;
  bool on;
  bool enabled;  // lanes >= `lanes per warp` never own an instance
  int LW;        // instances per warp = interleave stride of the workspace
  double* ws;    // this lane's column of the interleaved workspace
  // element offsets
  size_t VS, o_ri, o_gm, o_K, o_L, o_D, o_Di, o_y, o_x, o_Hx, o_f, o_Gx, o_h, o_Ax, o_b;
  const double *Hx, *f, *Gx, *h, *Ax, *b;  // this lane's instance, wire format

  __device__ __forceinline__ double W(size_t e) const { return ws[e * LW]; }
  __device__ __forceinline__ void S(size_t e, double v) const {
    if (on) ws[e * LW] = v;
  }
  // entry i of part (0 z, 1 l, 2 v, 3 y) of iterate block `blk`
  __device__ __forceinline__ size_t vz(int blk, int i) const { return blk * VS + i; }
  __device__ __forceinline__ size_t vl(int blk, int i) const { return blk * VS + nz + i; }
  __device__ __forceinline__ size_t vv(int blk, int i) const { return blk * VS + nz + nl + i; }
  __device__ __forceinline__ size_t vy(int blk, int i) const {
    return blk * VS + nz + nl + nv + i;
  }

  __device__ void layout(const SparseDev& dd) {
    d = dd;
    nz = d.nz;
    nl = d.nl;
    nv = d.nv;
    n = d.n;
    VS = (size_t)nz + nl + 2 * (size_t)nv;
    o_ri = 3 * VS;
    o_gm = o_ri + nz + nl + nv;
    o_K = o_gm + 3 * (size_t)nv;
    o_L = o_K + d.nnzK;
    o_D = o_L + d.nnzL;
    o_Di = o_D + n;
    o_y = o_Di + n;
    o_x = o_y + n;
    o_Hx = o_x + n;
    o_f = o_Hx + d.nnzH;
    o_Gx = o_f + nz;
    o_h = o_Gx + d.nnzG;
    o_Ax = o_h + nl;
    o_b = o_Ax + d.nnzA;
  }

  __device__ void bind(const SparseArgs& a, int inst) {
    const size_t i = (size_t)inst;
    Hx = a.Hx + i * d.nnzH;
    f = a.f + i * nz;
    Gx = a.Gx + i * d.nnzG;
    h = a.h + i * nl;
    Ax = a.Ax + i * d.nnzA;
    b = a.b + i * nv;
  }

  // CopyIntoVariable + InitializeConstraintMargin + xi = xk; returns the forcing norm
  // sqrt(b'b + f'f + h'h) (dense_data.h:72-73)
  __device__ double init(const double* z0, const double* l0, const double* v0) {
    on = true;
    // problem data: instance-major wire format -> this lane's workspace column
    for (int e = 0; e < d.nnzH; e++) S(o_Hx + e, __ldg(Hx + e));
    for (int e = 0; e < d.nnzG; e++) S(o_Gx + e, __ldg(Gx + e));
    for (int e = 0; e < d.nnzA; e++) S(o_Ax + e, __ldg(Ax + e));
    double sb = 0.0, sf = 0.0, sh = 0.0;
    for (int i = 0; i < nz; i++) {
      const double fv = __ldg(f + i);
      sf = fma(fv, fv, sf);
      S(o_f + i, fv);
      const double zv = z0[i];
      S(vz(O_XK, i), zv);
      S(vz(O_XI, i), zv);
    }
    for (int i = 0; i < nl; i++) {
      const double hv = __ldg(h + i);
      sh = fma(hv, hv, sh);
      S(o_h + i, hv);
      const double lv = l0[i];
      S(vl(O_XK, i), lv);
      S(vl(O_XI, i), lv);
    }
    for (int k = 0; k < nv; k++) {
      const double bv = __ldg(b + k);
      sb = fma(bv, bv, sb);
      S(o_b + k, bv);
      double s = 0.0;
      for (int q = d.Ar_ptr[k]; q < d.Ar_ptr[k + 1]; q++)
        s = fma(W(o_Ax + d.Ar_val[q]), W(vz(O_XK, d.Ar_col[q])), s);
      const double yv = bv - s;  // y = b - A z0 (full_variable.cc:47-53)
      const double vv0 = v0[k];
      S(vv(O_XK, k), vv0);
      S(vv(O_XI, k), vv0);
      S(vy(O_XK, k), yv);
      S(vy(O_XI, k), yv);
    }
    return sqrt((sb + sf) + sh);
  }

  // Fused residual evaluation at x = block `base` (+ t dx when trial): the inner
  // residual wrt xbar = xk (wrt x itself when self_bar) -> ri, and both norms
  // (engine.cuh::evaluate; full_residual.cc:49-109).  The trial point is formed on
  // the fly with the fused multiply-adds commit() uses.
  __device__ EvalOut evaluate(int base, bool trial, double t, bool self_bar, double sigma,
                              double alpha) {
    const double sg = self_bar ? 0.0 : sigma;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    auto zt = [&](int j) {
      const double x = W(vz(base, j));
      return trial ? fma(t, W(vz(O_DX, j)), x) : x;
    };
    auto lt = [&](int j) {
      const double x = W(vl(base, j));
      return trial ? fma(t, W(vl(O_DX, j)), x) : x;
    };
    auto vt = [&](int j) {
      const double x = W(vv(base, j));
      return trial ? fma(t, W(vv(O_DX, j)), x) : x;
    };
    // z block: tz = ((f + Hz) + G'l) + A'v
    for (int i = 0; i < nz; i++) {
      double s = 0.0, g = 0.0, a = 0.0;
      for (int q = d.Hr_ptr[i]; q < d.Hr_ptr[i + 1]; q++)
        s = fma(W(o_Hx + d.Hr_val[q]), zt(d.Hr_col[q]), s);
      for (int e = d.Gp[i]; e < d.Gp[i + 1]; e++) g = fma(W(o_Gx + e), lt(d.Gi[e]), g);
      for (int e = d.Ap[i]; e < d.Ap[i + 1]; e++) a = fma(W(o_Ax + e), vt(d.Ai[e]), a);
      const double tz = ((W(o_f + i) + s) + g) + a;
      s3 = fma(tz, tz, s3);
      const double r = tz + sg * (zt(i) - W(vz(O_XK, i)));
      S(o_ri + i, r);
      s0 = fma(r, r, s0);
    }
    // l block: tl = h - Gz
    for (int r_ = 0; r_ < nl; r_++) {
      double s = 0.0;
      for (int q = d.Gr_ptr[r_]; q < d.Gr_ptr[r_ + 1]; q++)
        s = fma(W(o_Gx + d.Gr_val[q]), zt(d.Gr_col[q]), s);
      const double tl = W(o_h + r_) - s;
      s4 = fma(tl, tl, s4);
      const double r = tl + sg * (lt(r_) - W(vl(O_XK, r_)));
      S(o_ri + nz + r_, r);
      s1 = fma(r, r, s1);
    }
    // v block
    for (int k = 0; k < nv; k++) {
      const double v = vt(k);
      double y = W(vy(base, k));
      if (trial) {  // y-aware axpy, full_variable.cc:55-65
        const double y1 = fma(t, W(vy(O_DX, k)), y);
        y = fma(-t, W(o_b + k), y1);
      }
      const double ys = y + sg * (v - W(vv(O_XK, k)));
      const double rv = pfb(ys, v, alpha);
      S(o_ri + nz + nl + k, rv);
      s2 = fma(rv, rv, s2);
      const double nn = pnr(y, v, alpha);
      s5 = fma(nn, nn, s5);
    }
    EvalOut e;
    const double zn = sqrt(s0), ln = sqrt(s1), vn = sqrt(s2);
    e.Ei = sqrt(zn * zn + ln * ln + vn * vn);
    const double zo = sqrt(s3), lo = sqrt(s4), vo = sqrt(s5);
    e.Eo = sqrt(zo * zo + lo * lo + vo * vo);
    return e;
  }

  // xi <- xi + t dx for the lanes with `c` (y-aware)
  __device__ void commit_if(bool c, double t) {
    const bool keep = on;
    on = keep && c;
    for (int i = 0; i < nz + nl + nv; i++)  // z, l, v are contiguous in a block
      S(vz(O_XI, i), fma(t, W(vz(O_DX, i)), W(vz(O_XI, i))));
    for (int k = 0; k < nv; k++) {
      const double y1 = fma(t, W(vy(O_DX, k)), W(vy(O_XI, k)));
      S(vy(O_XI, k), fma(-t, W(o_b + k), y1));
    }
    on = keep;
  }
  __device__ void commit(double t) { commit_if(true, t); }

  // LinearSolver::Initialize: barrier terms, K, numeric LDL' (the schedule of
  // QDLDL_factor, tools/qdldl/qdldl_wrapper.h:46-54).  False on a zero / NaN pivot.
  __device__ bool factor(double sigma, double alpha, bool with_commit, double t,
                         bool any_commit) {
    if (any_commit) commit_if(with_commit, t);
    for (int k = 0; k < nv; k++) {
      const double v = W(vv(O_XI, k));
      const double ys = W(vy(O_XI, k)) + sigma * (v - W(vv(O_XK, k)));
      double ga, mu;
      pfb_barrier(ys, v, alpha, sigma, &ga, &mu);
      S(o_gm + k, ga);
      S(o_gm + nv + k, mu);
      S(o_gm + 2 * nv + k, sqrt(div_nr(ga, mu)));
    }
    for (int e = 0; e < d.nnzK; e++) {
      const int kind = d.Kkind[e], idx = d.Kidx[e];
      double v;
      if (kind == KSRC_H) v = W(o_Hx + idx);
      else if (kind == KSRC_H_SIGMA) v = W(o_Hx + idx) + sigma;
      else if (kind == KSRC_SIGMA) v = sigma;
      else if (kind == KSRC_G) v = W(o_Gx + idx);
      else if (kind == KSRC_NEG_SIGMA) v = -sigma;
      else if (kind == KSRC_A) v = W(o_gm + 2 * nv + d.Krow[e]) * W(o_Ax + idx);
      else v = -1.0;
      S(o_K + e, v);
    }
    bool ok = true;
    for (int i = 0; i < n; i++) S(o_y + i, 0.0);
    for (int k = 0; k < n; k++) {
      double dk = 0.0;
      for (int p = d.Kp[k]; p < d.Kp[k + 1]; p++) {
        const int r = d.Ki[p];
        const double kv = W(o_K + p);
        if (r == k) dk = kv;
        else S(o_y + r, kv);
      }
      for (int q = d.Sp[k]; q < d.Sp[k + 1]; q++) {
        const int c = d.Sc[q], slot = d.St[q];
        const double yc = W(o_y + c);
        // (the rows of a column are distinct: four updates in flight at a time -- the
        // compiler cannot know that the stores do not alias the following loads)
        int j = d.Lp[c];
        for (; j + 4 <= slot; j += 4) {
          const size_t y0 = o_y + d.Li[j], y1 = o_y + d.Li[j + 1], y2 = o_y + d.Li[j + 2],
                       y3 = o_y + d.Li[j + 3];
          const double l0 = W(o_L + j), l1 = W(o_L + j + 1), l2 = W(o_L + j + 2),
                       l3 = W(o_L + j + 3);
          const double v0 = W(y0), v1 = W(y1), v2 = W(y2), v3 = W(y3);
          S(y0, fma(-l0, yc, v0));
          S(y1, fma(-l1, yc, v1));
          S(y2, fma(-l2, yc, v2));
          S(y3, fma(-l3, yc, v3));
        }
        for (; j < slot; j++) {
          const size_t yi = o_y + d.Li[j];
          S(yi, fma(-W(o_L + j), yc, W(yi)));
        }
        const double lx = yc * W(o_Di + c);
        S(o_L + slot, lx);
        dk = fma(-yc, lx, dk);
        S(o_y + c, 0.0);
      }
      if (!(fabs(dk) > 0.0)) ok = false;
      S(o_D + k, dk);
      S(o_Di + k, 1.0 / dk);
    }
    return ok;
  }

  // LinearSolver::Solve on r = -(ri) -> dx (QDLDL_solve: L, D^-1, L')
  __device__ void solve() {
    // r3 = rv ./ mu -> dx.v (scratch until dv is written)
    for (int k = 0; k < nv; k++)
      S(vv(O_DX, k), div_nr(-W(o_ri + nz + nl + k), W(o_gm + nv + k)));
    for (int i = 0; i < nz; i++) {
      double s = 0.0;
      for (int e = d.Ap[i]; e < d.Ap[i + 1]; e++) s = fma(W(o_Ax + e), W(vv(O_DX, d.Ai[e])), s);
      S(o_x + d.iperm[i], (-W(o_ri + i)) - s);
    }
    for (int r = 0; r < nl; r++) S(o_x + d.iperm[nz + r], W(o_ri + nz + r));
    for (int k = 0; k < nv; k++) S(o_x + d.iperm[nz + nl + k], 0.0);
    for (int i = 0; i < n; i++) {
      const double xi = W(o_x + i);
      const int je = d.Lp[i + 1];
      int j = d.Lp[i];
      for (; j + 4 <= je; j += 4) {
        const size_t t0 = o_x + d.Li[j], t1 = o_x + d.Li[j + 1], t2 = o_x + d.Li[j + 2],
                     t3 = o_x + d.Li[j + 3];
        const double l0 = W(o_L + j), l1 = W(o_L + j + 1), l2 = W(o_L + j + 2),
                     l3 = W(o_L + j + 3);
        const double v0 = W(t0), v1 = W(t1), v2 = W(t2), v3 = W(t3);
        S(t0, fma(-l0, xi, v0));
        S(t1, fma(-l1, xi, v1));
        S(t2, fma(-l2, xi, v2));
        S(t3, fma(-l3, xi, v3));
      }
      for (; j < je; j++) {
        const size_t t = o_x + d.Li[j];
        S(t, fma(-W(o_L + j), xi, W(t)));
      }
    }
    for (int i = 0; i < n; i++) S(o_x + i, W(o_x + i) * W(o_Di + i));
    for (int i = n - 1; i >= 0; i--) {
      double xi = W(o_x + i);
      for (int j = d.Lp[i]; j < d.Lp[i + 1]; j++) xi = fma(-W(o_L + j), W(o_x + d.Li[j]), xi);
      S(o_x + i, xi);
    }
    for (int i = 0; i < nz; i++) S(vz(O_DX, i), W(o_x + d.iperm[i]));
    for (int r = 0; r < nl; r++) S(vl(O_DX, r), W(o_x + d.iperm[nz + r]));
    // dv = (rv + gamma .* (A dz)) ./ mu ; dy = b - A dz
    for (int k = 0; k < nv; k++) {
      double s = 0.0;
      for (int q = d.Ar_ptr[k]; q < d.Ar_ptr[k + 1]; q++)
        s = fma(W(o_Ax + d.Ar_val[q]), W(vz(O_DX, d.Ar_col[q])), s);
      S(vv(O_DX, k),
        div_nr(W(o_gm + k) * s + (-W(o_ri + nz + nl + k)), W(o_gm + nv + k)));
      S(vy(O_DX, k), W(o_b + k) - s);
    }
  }

  // End of a proximal subproblem (impl:301, 202-216): [commit,] ProjectDuals on xi;
  // for the lanes with do_diff: dx = xi - xk (y-aware), its norm, CheckFeasibility on
  // it (full_feasibility.cc:25-88; status in *feas) and xk <- xi.
  __device__ double prox_end(bool do_diff, bool with_commit, double t, bool any_commit,
                             double tol, bool check, int* feas) {
    const bool lanes = on;
    if (any_commit) commit_if(with_commit, t);
    for (int k = 0; k < nv; k++) S(vv(O_XI, k), fmax(W(vv(O_XI, k)), 0.0));
    on = lanes && do_diff;
    double sz = 0, sl = 0, sv = 0;
    for (int i = 0; i < nz; i++) {
      const double xi = W(vz(O_XI, i));
      const double dz = xi + (-1.0) * W(vz(O_XK, i));
      sz = fma(dz, dz, sz);
      S(vz(O_DX, i), dz);
      S(vz(O_XK, i), xi);
    }
    for (int i = 0; i < nl; i++) {
      const double xi = W(vl(O_XI, i));
      const double dl = xi + (-1.0) * W(vl(O_XK, i));
      sl = fma(dl, dl, sl);
      S(vl(O_DX, i), dl);
      S(vl(O_XK, i), xi);
    }
    for (int k = 0; k < nv; k++) {
      const double xi = W(vv(O_XI, k)), yi = W(vy(O_XI, k));
      const double dv = xi + (-1.0) * W(vv(O_XK, k));
      sv = fma(dv, dv, sv);
      const double dy = yi + (-1.0) * W(vy(O_XK, k));
      S(vv(O_DX, k), dv);
      S(vy(O_DX, k), dy + W(o_b + k));  // y-aware: y += (-a) b with a = -1
      S(vv(O_XK, k), xi);
      S(vy(O_XK, k), yi);
    }
    *feas = 0;
    if (check) {
      double d1 = -INFINITY, d2 = 0, d3 = 0, d4 = 0, w = 0, p1 = 0, p2 = 0, umax = 0;
      for (int k = 0; k < nv; k++) {
        double s = 0.0;
        for (int q = d.Ar_ptr[k]; q < d.Ar_ptr[k + 1]; q++)
          s = fma(W(o_Ax + d.Ar_val[q]), W(vz(O_DX, d.Ar_col[q])), s);
        d1 = fmax(d1, s);
        const double dv = W(vv(O_DX, k));
        umax = fmax(umax, fabs(dv));
        p2 = fma(W(o_b + k), dv, p2);
      }
      for (int r = 0; r < nl; r++) {
        double s = 0.0;
        for (int q = d.Gr_ptr[r]; q < d.Gr_ptr[r + 1]; q++)
          s = fma(W(o_Gx + d.Gr_val[q]), W(vz(O_DX, d.Gr_col[q])), s);
        d2 = fmax(d2, fabs(s));
        const double dl = W(vl(O_DX, r));
        umax = fmax(umax, fabs(dl));
        p2 = fma(W(o_h + r), dl, p2);
      }
      for (int i = 0; i < nz; i++) {
        double s = 0.0, a = 0.0, g = 0.0;
        for (int q = d.Hr_ptr[i]; q < d.Hr_ptr[i + 1]; q++)
          s = fma(W(o_Hx + d.Hr_val[q]), W(vz(O_DX, d.Hr_col[q])), s);
        d3 = fmax(d3, fabs(s));
        const double dz = W(vz(O_DX, i));
        w = fmax(w, fabs(dz));
        d4 = fma(W(o_f + i), dz, d4);
        for (int e = d.Ap[i]; e < d.Ap[i + 1]; e++) a = fma(W(o_Ax + e), W(vv(O_DX, d.Ai[e])), a);
        for (int e = d.Gp[i]; e < d.Gp[i + 1]; e++) g = fma(W(o_Gx + e), W(vl(O_DX, d.Gi[e])), g);
        p1 = fmax(p1, fabs(a + g));
      }
      const bool dual_inf = (d1 <= w * tol) && (d2 <= tol * w) && (d3 <= tol * w) &&
                            (d4 < 0.0) && (w > 1e-14);
      const bool primal_inf = (p1 <= tol * umax) && (p2 < 0.0);
      *feas = (primal_inf ? 1 : 0) + (dual_inf ? 2 : 0);
    }
    on = lanes;
    const double a = sqrt(sz), b2 = sqrt(sl), c = sqrt(sv);
    return sqrt(a * a + b2 * b2 + c * c);
  }

  __device__ void write_result(int from, double* z, double* l, double* v, double* y) {
    for (int i = 0; i < nz; i++) z[i] = W(vz(from, i));
    for (int i = 0; i < nl; i++) l[i] = W(vl(from, i));
    for (int k = 0; k < nv; k++) {
      v[k] = W(vv(from, k));
      y[k] = W(vy(from, k));
    }
  }
};

constexpr int kSparseWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kSparseWarpsPerCta, 2)
sparse_lane_kernel(const __grid_constant__ SparseArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= a.warps) return;
  SparseLane p;
  p.layout(a.d);
  p.LW = a.lanes;
  p.enabled = lane < a.lanes;
  // (idle lanes alias lane 0's column: they execute the warp's sweeps with their stores off)
  p.ws = a.ws + (size_t)warp * a.ws_stride + (p.enabled ? lane : 0);
  p.on = true;
  p.bind(a, 0);
  lane_solve_loop(p, a);
}

}  // namespace

// Instances per warp.  Measured on the servo OCP as a sparse QP (16,384 instances,
// profiles/r2_sparse_lane.txt): 32 lanes 15.1 k, 16 lanes 14.6 k, 8 lanes 12.7 k solves/s --
// the time is the dependent chain of indirectly addressed loads of ONE warp times the
// rounds of its slowest instance, which more (narrower) warps do not shorten.  Full warps
// are the default; FBSTAB_SPARSE_LANES = 16 / 8 re-measures.
int SparseLaneLanes(int batch, int sms) {
  (void)batch;
  (void)sms;
  if (const char* e = getenv("FBSTAB_SPARSE_LANES")) {
    const int v = atoi(e);
    if (v == 8 || v == 16 || v == 32) return v;
  }
  return 32;
}

size_t SparseLaneWsDoublesPerLane(const SparseDev& d) {
  const size_t vs = (size_t)d.nz + d.nl + 2 * (size_t)d.nv;
  const size_t per_lane = 3 * vs + ((size_t)d.nz + d.nl + d.nv) + 3 * (size_t)d.nv + d.nnzK +
                          d.nnzL + 4 * (size_t)d.n + d.nnzH + d.nz + d.nnzG + d.nl + d.nnzA + d.nv;
  return per_lane;
}

// Instances resident at a time (each needs a workspace column) for batches up to `batch`.
int SparseLaneSlots(int batch, int sms) {
  const int lanes = SparseLaneLanes(batch, sms);
  const int need = std::max(1, (batch + lanes - 1) / lanes);
  return lanes * std::min(need, sms * 4 * kSparseWarpsPerCta);
}

int SparseLaneLaunch(const SparseDev& d, int batch, int lane_slots, const double* Hx, const double* f,
                     const double* Gx, const double* h, const double* Ax, const double* b,
                     double* z, double* l, double* v, double* y, fbstab_out* out,
                     const fbstab_options& opts, double* ws, int* counter, cudaStream_t stream) {
  SparseArgs a;
  a.d = d;
  a.batch = batch;
  a.Hx = Hx;
  a.f = f;
  a.Gx = Gx;
  a.h = h;
  a.Ax = Ax;
  a.b = b;
  a.z = z;
  a.l = l;
  a.v = v;
  a.y = y;
  a.out = out;
  a.ws = ws;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  a.lanes = SparseLaneLanes(batch, sms);
  a.ws_stride = (size_t)a.lanes * SparseLaneWsDoublesPerLane(d);
  a.counter = counter;
  a.opts = opts;
  a.warps = std::max(1, std::min(lane_slots / a.lanes, (batch + a.lanes - 1) / a.lanes));
  // spread the warps over the SMs: CTAs of up to kSparseWarpsPerCta warps
  const int per_cta = std::max(1, std::min(kSparseWarpsPerCta, (a.warps + 4 * sms - 1) / (4 * sms)));
  const int ctas = (a.warps + per_cta - 1) / per_cta;
  // warp index = blockIdx.x * warps per CTA + warp in CTA
  sparse_lane_kernel<<<ctas, 32 * per_cta, 0, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

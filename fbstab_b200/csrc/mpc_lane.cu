// mpc_lane.cu -- lane-per-instance FBstab for MPC problems with SMALL stages
// (nx <= 4: the servo-motor and double-integrator OCPs of BASELINE config 3).
//
// Why a second MPC kernel.  The CTA-per-instance kernel (mpc_riccati.cu) is
// instruction bound at these sizes: the Riccati chain works on 2x2..4x4 blocks,
// so at most 5 of a warp's 32 lanes do useful work and a solve costs 11 M warp
// instructions (profiles/r1_mpc_generic_instruction_bound.txt).  Here every
// LANE owns one instance: all stage matrices live in registers, every loop is
// unrolled at compile time, there is no shuffle, no barrier and no idle lane,
// and one warp instruction serves 32 instances.
//
//  * Per-instance state (iterates, residual, step, barrier terms, the factor
//    blocks of all N+1 stages) lives in a per-warp global workspace that is
//    LANE-INTERLEAVED: element e of lane j sits at ws[e*32 + j], so every
//    state access of the warp is one coalesced 256-byte transaction.
//  * The problem data is read in place from the reference's instance-major
//    wire format; a lane streams its own sequences front to back, so each
//    32-byte sector it pulls is fully used from L1 by the following reads.
//  * The recursion, its operation order and the fused residual evaluation are
//    those of mpc_riccati.cuh / engine.cuh (reference
//    riccati_linear_solver.cc:77-344, mpc_data.cc:17-289,
//    fbstab_algorithm-impl.h:113-304); the algorithm is a per-lane phase
//    machine, and the warp runs "evaluate -> decide -> Newton step -> end of
//    subproblem" rounds in which a lane takes part in the stages it needs.
//  * Lanes pull instance indices from the global atomic counter on their own,
//    so a lane whose instance has converged starts the next one in the
//    following round (no host round trips, no waiting for the slowest lane).
//  * The trial point x + t dx of the Armijo search is never materialised: the
//    evaluation forms it on the fly and an accepted step is committed with
//    the same fused multiply-add, bit for bit.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.cuh"
#include "mpc_lane.h"

namespace fbs {
namespace {

struct LaneArgs {
  int N, batch;
  MpcData data;
  double *z, *l, *v, *y;
  fbstab_out* out;
  double* ws;        // per-warp workspace base
  size_t ws_stride;  // doubles per warp
  int* counter;
  fbstab_options opts;
  int comp;
  fbstab_component_io io;
};

enum { PH_TOP = 0, PH_TRIAL = 1, PH_REEVAL = 2, PH_FINAL = 3 };

// ---- register-resident small dense algebra (operation order of the team
// versions in mpc_riccati.cuh) ------------------------------------------------
template <int M>
__device__ __forceinline__ bool chol(double (&A)[M][M]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < M; k++) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < k; j++) s = fma(A[k][j], A[k][j], s);
    double x = A[k][k] - s;
    if (!(x > 0.0)) ok = false;
    x = sqrt(x);
#pragma unroll
    for (int i = k + 1; i < M; i++) {
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < k; j++) a = fma(A[i][j], A[k][j], a);
      A[i][k] = (A[i][k] - a) / x;
    }
    A[k][k] = x;
  }
  return ok;
}
// y = L^-1 x (x destroyed)
template <int M>
__device__ __forceinline__ void trsv_l(const double (&L)[M][M], double (&x)[M], double (&y)[M]) {
#pragma unroll
  for (int j = 0; j < M; j++) {
    const double xj = x[j] / L[j][j];
    y[j] = xj;
#pragma unroll
    for (int i = j + 1; i < M; i++) x[i] = fma(-L[i][j], xj, x[i]);
  }
}
// y = L^-T x (x destroyed)
template <int M>
__device__ __forceinline__ void trsv_lt(const double (&L)[M][M], double (&x)[M], double (&y)[M]) {
#pragma unroll
  for (int i = M - 1; i >= 0; i--) {
    const double xi = x[i] / L[i][i];
    y[i] = xi;
#pragma unroll
    for (int r = 0; r < i; r++) x[r] = fma(-L[i][r], xi, x[r]);
  }
}
// X(row) = src(row) L^-T
template <int M>
__device__ __forceinline__ void row_trsm_lt(const double (&L)[M][M], const double (&src)[M],
                                            double (&X)[M]) {
#pragma unroll
  for (int j = 0; j < M; j++) {
    double s = src[j];
#pragma unroll
    for (int k = 0; k < j; k++) s = fma(-X[k], L[j][k], s);
    X[j] = s / L[j][j];
  }
}

template <int NX, int NU, int NC>
struct Lane {
  static constexpr int NS = NX + NU;
  static constexpr int TX = NX * (NX + 1) / 2, TU = NU * (NU + 1) / 2;
  // factor block of a stage: lower L | lower M | AM | SM | P | lower SG
  static constexpr int oL = 0, oM = TX, oAM = 2 * TX, oSM = oAM + NX * NX,
                       oP = oSM + NU * NX, oSG = oP + NX * NU, FS = oSG + TU;

  int N, nz, nl, nv;
  double* ws;  // this lane's column of the interleaved workspace
  int o_xk, o_xi, o_dx, o_ri, o_gm, o_fac;
  const double *Q, *R, *S, *q, *r, *A, *B, *c, *E, *L, *d, *x0;

  __device__ __forceinline__ double ld(int e) const { return ws[(size_t)e * 32]; }
  __device__ __forceinline__ void st(int e, double v) const { ws[(size_t)e * 32] = v; }
  // Vars block layout: z | l | v | y
  __device__ __forceinline__ int zo(int base, int i) const { return base + i * NS; }
  __device__ __forceinline__ int lo(int base, int i) const { return base + nz + i * NX; }
  __device__ __forceinline__ int vo(int base, int i) const { return base + nz + nl + i * NC; }
  __device__ __forceinline__ int yo(int base, int i) const { return base + nz + nl + nv + i * NC; }

  __device__ void bind(const LaneArgs& a, int inst) {
    const size_t i = (size_t)inst, K = N + 1;
    Q = a.data.Q + i * K * NX * NX;
    R = a.data.R + i * K * NU * NU;
    S = a.data.S + i * K * NU * NX;
    q = a.data.q + i * K * NX;
    r = a.data.r + i * K * NU;
    A = a.data.A + i * N * NX * NX;
    B = a.data.B + i * N * NX * NU;
    c = a.data.c + i * N * NX;
    E = a.data.E + i * K * NC * NX;
    L = a.data.L + i * K * NC * NU;
    d = a.data.d + i * K * NC;
    x0 = a.data.x0 + i * NX;
  }

  __device__ __forceinline__ void load_lower(int e, double (&M)[NX][NX]) const {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NX; cc++)
#pragma unroll
      for (int rr = cc; rr < NX; rr++) M[rr][cc] = ld(e + k++);
  }
  __device__ __forceinline__ void store_lower(int e, const double (&M)[NX][NX]) const {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NX; cc++)
#pragma unroll
      for (int rr = cc; rr < NX; rr++) st(e + k++, M[rr][cc]);
  }
  __device__ __forceinline__ void load_lower_u(int e, double (&M)[NU][NU]) const {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NU; cc++)
#pragma unroll
      for (int rr = cc; rr < NU; rr++) M[rr][cc] = ld(e + k++);
  }
  __device__ __forceinline__ void store_lower_u(int e, const double (&M)[NU][NU]) const {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NU; cc++)
#pragma unroll
      for (int rr = cc; rr < NU; rr++) st(e + k++, M[rr][cc]);
  }

  // (E(i) x + L(i) u)[k], mpc_data.cc:66-105
  __device__ __forceinline__ double Az_entry(const double (&z)[NS], int i, int k) const {
    const double* Em = E + (size_t)i * NC * NX;
    const double* Lm = L + (size_t)i * NC * NU;
    double s = 0.0, s2 = 0.0;
#pragma unroll
    for (int cc = 0; cc < NX; cc++) s = fma(__ldg(Em + k + cc * NC), z[cc], s);
#pragma unroll
    for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Lm + k + cc * NC), z[NX + cc], s2);
    return s + s2;
  }

  // CopyIntoVariable + InitializeConstraintMargin + xi = xk; returns the forcing norm
  __device__ double init(const double* z0, const double* l0, const double* v0) {
    double fn = 0.0;
    for (int i = 0; i <= N; i++) {
      double z[NS];
#pragma unroll
      for (int k = 0; k < NS; k++) {
        z[k] = z0[i * NS + k];
        st(zo(o_xk, i) + k, z[k]);
        st(zo(o_xi, i) + k, z[k]);
      }
#pragma unroll
      for (int k = 0; k < NX; k++) {
        const double lv = l0[i * NX + k];
        st(lo(o_xk, i) + k, lv);
        st(lo(o_xi, i) + k, lv);
        const double qv = __ldg(q + i * NX + k);
        fn = fma(qv, qv, fn);
        if (i < N) {
          const double cv = __ldg(c + i * NX + k);
          fn = fma(cv, cv, fn);
        }
      }
#pragma unroll
      for (int k = 0; k < NU; k++) {
        const double rv = __ldg(r + i * NU + k);
        fn = fma(rv, rv, fn);
      }
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const double vv = v0[i * NC + k];
        const double dv = __ldg(d + i * NC + k);
        fn = fma(dv, dv, fn);
        const double yv = -dv - Az_entry(z, i, k);
        st(vo(o_xk, i) + k, vv);
        st(vo(o_xi, i) + k, vv);
        st(yo(o_xk, i) + k, yv);
        st(yo(o_xi, i) + k, yv);
      }
    }
#pragma unroll
    for (int k = 0; k < NX; k++) fn = fma(__ldg(x0 + k), __ldg(x0 + k), fn);
    return sqrt(fn);
  }

  // Fused residual evaluation at x = base (+ t dx when trial): inner residual
  // wrt xbar = xk (or wrt x itself when self_bar) -> ri, both norms.
  __device__ EvalOut evaluate(int base, bool trial, double t, bool self_bar, double sigma,
                              double alpha) {
    const double sg = self_bar ? 0.0 : sigma;
    double s[6] = {0, 0, 0, 0, 0, 0};
    double zp[NS], lc[NX], ln[NX];  // z(i-1), l(i), l(i+1)
    auto get = [&](int off_base, int off_dx) {
      double v = ld(off_base);
      if (trial) v = fma(t, ld(off_dx), v);
      return v;
    };
#pragma unroll
    for (int k = 0; k < NX; k++) lc[k] = get(lo(base, 0) + k, lo(o_dx, 0) + k);
#pragma unroll
    for (int k = 0; k < NS; k++) zp[k] = 0.0;
    for (int i = 0; i <= N; i++) {
      double z[NS], v[NC], y[NC];
#pragma unroll
      for (int k = 0; k < NS; k++) z[k] = get(zo(base, i) + k, zo(o_dx, i) + k);
#pragma unroll
      for (int k = 0; k < NC; k++) {
        v[k] = get(vo(base, i) + k, vo(o_dx, i) + k);
        double yv = ld(yo(base, i) + k);
        if (trial) {  // y-aware axpy, full_variable.cc:55-65
          yv = fma(t, ld(yo(o_dx, i) + k), yv);
          yv = fma(-t, -__ldg(d + i * NC + k), yv);
        }
        y[k] = yv;
      }
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++) ln[k] = get(lo(base, i + 1) + k, lo(o_dx, i + 1) + k);
      }
      const double* Qm = Q + (size_t)i * NX * NX;
      const double* Rm = R + (size_t)i * NU * NU;
      const double* Sm = S + (size_t)i * NU * NX;
      const double* Em = E + (size_t)i * NC * NX;
      const double* Lm = L + (size_t)i * NC * NU;
      const double* Am = A + (size_t)i * NX * NX;
      const double* Bm = B + (size_t)i * NX * NU;
      // z block: tz = ((f + Hz) + G'l) + A'v
#pragma unroll
      for (int rr = 0; rr < NS; rr++) {
        double s1 = 0.0, s2 = 0.0, tz;
        if (rr < NX) {
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s1 = fma(__ldg(Qm + rr + cc * NX), z[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Sm + cc + rr * NU), z[NX + cc], s2);
          tz = __ldg(q + i * NX + rr) + (s1 + s2);
          tz += -lc[rr];
          if (i < N) {
            double sa = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) sa = fma(__ldg(Am + cc + rr * NX), ln[cc], sa);
            tz += sa;
          }
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(__ldg(Em + k + rr * NC), v[k], sv);
          tz += sv;
        } else {
          const int ru = rr - NX;
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s1 = fma(__ldg(Sm + ru + cc * NU), z[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Rm + ru + cc * NU), z[NX + cc], s2);
          tz = __ldg(r + i * NU + ru) + (s1 + s2);
          if (i < N) {
            double sa = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) sa = fma(__ldg(Bm + cc + ru * NX), ln[cc], sa);
            tz += sa;
          }
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(__ldg(Lm + k + ru * NC), v[k], sv);
          tz += sv;
        }
        s[3] = fma(tz, tz, s[3]);
        const double rr_ = tz + sg * (z[rr] - ld(zo(o_xk, i) + rr));
        st(o_ri + i * NS + rr, rr_);
        s[0] = fma(rr_, rr_, s[0]);
      }
      // l block: tl = h - Gz
#pragma unroll
      for (int rr = 0; rr < NX; rr++) {
        double tl;
        if (i == 0) {
          tl = -__ldg(x0 + rr) + z[rr];
        } else {
          const double* Ap = A + (size_t)(i - 1) * NX * NX;
          const double* Bp = B + (size_t)(i - 1) * NX * NU;
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s1 = fma(__ldg(Ap + rr + cc * NX), zp[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Bp + rr + cc * NX), zp[NX + cc], s2);
          tl = (-__ldg(c + (size_t)(i - 1) * NX + rr) - (s1 + s2)) + z[rr];
        }
        s[4] = fma(tl, tl, s[4]);
        const double rl = tl + sg * (lc[rr] - ld(lo(o_xk, i) + rr));
        st(o_ri + nz + i * NX + rr, rl);
        s[1] = fma(rl, rl, s[1]);
      }
      // v block
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const double ys = y[k] + sg * (v[k] - ld(vo(o_xk, i) + k));
        const double rv = pfb(ys, v[k], alpha);
        st(o_ri + nz + nl + i * NC + k, rv);
        s[2] = fma(rv, rv, s[2]);
        const double nn = pnr(y[k], v[k], alpha);
        s[5] = fma(nn, nn, s[5]);
      }
#pragma unroll
      for (int k = 0; k < NS; k++) zp[k] = z[k];
#pragma unroll
      for (int k = 0; k < NX; k++) lc[k] = ln[k];
    }
    EvalOut e;
    const double zn = sqrt(s[0]), l_n = sqrt(s[1]), vn = sqrt(s[2]);
    e.Ei = sqrt(zn * zn + l_n * l_n + vn * vn);
    const double zz = sqrt(s[3]), lz = sqrt(s[4]), vz = sqrt(s[5]);
    e.Eo = sqrt(zz * zz + lz * lz + vz * vz);
    return e;
  }

  // xi <- xi + t dx (same fused multiply-adds as the trial evaluation)
  __device__ void commit(double t) {
    for (int i = 0; i <= N; i++) {
#pragma unroll
      for (int k = 0; k < NS; k++)
        st(zo(o_xi, i) + k, fma(t, ld(zo(o_dx, i) + k), ld(zo(o_xi, i) + k)));
#pragma unroll
      for (int k = 0; k < NX; k++)
        st(lo(o_xi, i) + k, fma(t, ld(lo(o_dx, i) + k), ld(lo(o_xi, i) + k)));
#pragma unroll
      for (int k = 0; k < NC; k++) {
        st(vo(o_xi, i) + k, fma(t, ld(vo(o_dx, i) + k), ld(vo(o_xi, i) + k)));
        double yv = fma(t, ld(yo(o_dx, i) + k), ld(yo(o_xi, i) + k));
        yv = fma(-t, -__ldg(d + i * NC + k), yv);
        st(yo(o_xi, i) + k, yv);
      }
    }
  }

  // RiccatiLinearSolver::Initialize, riccati_linear_solver.cc:77-210
  __device__ bool factor(double sigma, double alpha) {
    bool ok = true;
    double Lc[NX][NX];
    const double rs = sqrt(sigma);
#pragma unroll
    for (int a_ = 0; a_ < NX; a_++)
#pragma unroll
      for (int b_ = 0; b_ < NX; b_++) Lc[a_][b_] = (a_ == b_) ? rs : 0.0;
    for (int i = 0; i <= N; i++) {
      const int fb = o_fac + i * FS;
      store_lower(fb + oL, Lc);
      double Gam[NC];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const double yv = ld(yo(o_xi, i) + k), vv = ld(vo(o_xi, i) + k);
        const double ys = yv + sigma * (vv - ld(vo(o_xk, i) + k));
        double ga, mu;
        pfb_barrier(ys, vv, alpha, sigma, &ga, &mu);
        st(o_gm + i * NC + k, ga);
        st(o_gm + nv + i * NC + k, mu);
        Gam[k] = ga / mu;
      }
      const double* Qm = Q + (size_t)i * NX * NX;
      const double* Rm = R + (size_t)i * NU * NU;
      const double* Sm = S + (size_t)i * NU * NX;
      const double* Em = E + (size_t)i * NC * NX;
      const double* Lm = L + (size_t)i * NC * NU;
      double Ee[NC][NX], Le[NC][NU];
#pragma unroll
      for (int k = 0; k < NC; k++) {
#pragma unroll
        for (int cc = 0; cc < NX; cc++) Ee[k][cc] = __ldg(Em + k + cc * NC);
#pragma unroll
        for (int cc = 0; cc < NU; cc++) Le[k][cc] = __ldg(Lm + k + cc * NC);
      }
      // Linv = inv(L L'), column by column (:142-144)
      double Mm[NX][NX];
#pragma unroll
      for (int cc = 0; cc < NX; cc++) {
        double w[NX];
#pragma unroll
        for (int k = 0; k < NX; k++) w[k] = (k == cc) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < NX; j++) {
          w[j] /= Lc[j][j];
          const double wj = w[j];
#pragma unroll
          for (int k = j + 1; k < NX; k++) w[k] = fma(-Lc[k][j], wj, w[k]);
        }
#pragma unroll
        for (int k = NX - 1; k >= 0; k--) {
          double sv = w[k];
#pragma unroll
          for (int j = k + 1; j < NX; j++) sv = fma(-Lc[j][k], w[j], sv);
          w[k] = sv / Lc[k][k];
        }
#pragma unroll
        for (int rr = cc; rr < NX; rr++) Mm[rr][cc] = w[rr];
      }
      // M = chol(Q~ + Linv), Q~ = Q + sigma I + E' Gamma E (:102-123,145-147)
#pragma unroll
      for (int cc = 0; cc < NX; cc++)
#pragma unroll
        for (int rr = cc; rr < NX; rr++) {
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Ee[k][rr], Gam[k] * Ee[k][cc], sv);
          const double qt = (__ldg(Qm + rr + cc * NX) + (rr == cc ? sigma : 0.0)) + sv;
          Mm[rr][cc] = qt + Mm[rr][cc];
        }
      ok = chol<NX>(Mm) && ok;
      store_lower(fb + oM, Mm);
      // R~, S~
      double Rt[NU][NU], St[NU][NX];
#pragma unroll
      for (int cc = 0; cc < NU; cc++)
#pragma unroll
        for (int rr = cc; rr < NU; rr++) {
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Le[k][rr], Gam[k] * Le[k][cc], sv);
          Rt[rr][cc] = (__ldg(Rm + rr + cc * NU) + (rr == cc ? sigma : 0.0)) + sv;
        }
#pragma unroll
      for (int cc = 0; cc < NX; cc++)
#pragma unroll
        for (int rr = 0; rr < NU; rr++) {
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Le[k][rr], Gam[k] * Ee[k][cc], sv);
          St[rr][cc] = __ldg(Sm + rr + cc * NU) + sv;
        }
      // AM = A M^-T, SM = S~ M^-T (:149-161)
      double AM[NX][NX], SM[NU][NX];
      if (i < N) {
        const double* Am = A + (size_t)i * NX * NX;
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double src[NX];
#pragma unroll
          for (int cc = 0; cc < NX; cc++) src[cc] = __ldg(Am + rr + cc * NX);
          row_trsm_lt<NX>(Mm, src, AM[rr]);
        }
#pragma unroll
        for (int rr = 0; rr < NX; rr++)
#pragma unroll
          for (int cc = 0; cc < NX; cc++) st(fb + oAM + rr + cc * NX, AM[rr][cc]);
      }
#pragma unroll
      for (int rr = 0; rr < NU; rr++) row_trsm_lt<NX>(Mm, St[rr], SM[rr]);
#pragma unroll
      for (int rr = 0; rr < NU; rr++)
#pragma unroll
        for (int cc = 0; cc < NX; cc++) st(fb + oSM + rr + cc * NU, SM[rr][cc]);
      // SG = chol(R~ - SM SM') (:163-166)
      double SG[NU][NU];
#pragma unroll
      for (int cc = 0; cc < NU; cc++)
#pragma unroll
        for (int rr = 0; rr < NU; rr++) {
          double sv = 0.0;
          if (rr >= cc) {
#pragma unroll
            for (int k = 0; k < NX; k++) sv = fma(SM[rr][k], SM[cc][k], sv);
            sv = Rt[rr][cc] - sv;
          }
          SG[rr][cc] = sv;
        }
      ok = chol<NU>(SG) && ok;
      store_lower_u(fb + oSG, SG);
      if (i == N) break;
      // P = (AM SM' - B) SG^-T (:170-175)
      double P[NX][NU];
      {
        const double* Bm = B + (size_t)i * NX * NU;
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double src[NU];
#pragma unroll
          for (int j = 0; j < NU; j++) {
            double sv = 0.0;
#pragma unroll
            for (int k = 0; k < NX; k++) sv = fma(AM[rr][k], SM[j][k], sv);
            src[j] = sv - __ldg(Bm + rr + j * NX);
          }
          row_trsm_lt<NU>(SG, src, P[rr]);
#pragma unroll
          for (int j = 0; j < NU; j++) st(fb + oP + rr + j * NX, P[rr][j]);
        }
      }
      // L(i+1) = chol(sigma I + P P' + AM AM') (:179-183)
#pragma unroll
      for (int cc = 0; cc < NX; cc++)
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double v = 0.0;
          if (rr >= cc) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int k = 0; k < NU; k++) s1 = fma(P[rr][k], P[cc][k], s1);
#pragma unroll
            for (int k = 0; k < NX; k++) s2 = fma(AM[rr][k], AM[cc][k], s2);
            v = ((rr == cc ? sigma : 0.0) + s1) + s2;
          }
          Lc[rr][cc] = v;
        }
      ok = chol<NX>(Lc) && ok;
    }
    return ok;
  }

  // RiccatiLinearSolver::Solve on r = -(ri), riccati_linear_solver.cc:212-344.
  // Storage reuse as in mpc_riccati.cuh: theta(i) -> dx.l(i), M^-1 h -> dx.z x(i),
  // SG^-1(.) -> dx.z u(i) until the backward sweep writes the step there.
  __device__ void solve() {
    double th[NX];  // theta(i)
#pragma unroll
    for (int k = 0; k < NX; k++) th[k] = ld(o_ri + nz + k);  // r2(0) = rl(0)
    double lp[NX];  // dl(i+1) in the backward sweep
    for (int i = 0; i <= N; i++) {
      const int fb = o_fac + i * FS;
      // r3 = rv ./ mus, r1 = r.z - A' r3 for this stage (:222-225)
      double tv[NC], r1[NS];
#pragma unroll
      for (int k = 0; k < NC; k++)
        tv[k] = (-ld(o_ri + nz + nl + i * NC + k)) / ld(o_gm + nv + i * NC + k);
      {
        const double* Em = E + (size_t)i * NC * NX;
        const double* Lm = L + (size_t)i * NC * NU;
#pragma unroll
        for (int rr = 0; rr < NS; rr++) {
          double sv = 0.0;
          if (rr < NX) {
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(__ldg(Em + k + rr * NC), tv[k], sv);
          } else {
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(__ldg(Lm + k + (rr - NX) * NC), tv[k], sv);
          }
          r1[rr] = (-ld(o_ri + i * NS + rr)) - sv;
        }
      }
      double Lc[NX][NX], Mm[NX][NX], SG[NU][NU], SM[NU][NX];
      load_lower(fb + oL, Lc);
      load_lower(fb + oM, Mm);
      load_lower_u(fb + oSG, SG);
#pragma unroll
      for (int rr = 0; rr < NU; rr++)
#pragma unroll
        for (int cc = 0; cc < NX; cc++) SM[rr][cc] = ld(fb + oSM + rr + cc * NU);
      // h(i) = (L L')^-1 theta(i) - r1x(i)
      double sa[NX], sb[NX], sc[NX], tx[NX];
#pragma unroll
      for (int k = 0; k < NX; k++) {
        sa[k] = th[k];
        st(lo(o_dx, i) + k, th[k]);
      }
      trsv_l<NX>(Lc, sa, sb);
      trsv_lt<NX>(Lc, sb, sc);
#pragma unroll
      for (int k = 0; k < NX; k++) sa[k] = sc[k] - r1[k];
      trsv_l<NX>(Mm, sa, tx);  // tx = M^-1 h
      double ub[NU], tu[NU];
#pragma unroll
      for (int k = 0; k < NU; k++) {
        double sv = 0.0;
#pragma unroll
        for (int cc = 0; cc < NX; cc++) sv = fma(SM[k][cc], tx[cc], sv);
        ub[k] = sv + r1[NX + k];
      }
      if (i < N) {
        trsv_l<NU>(SG, ub, tu);  // tu = SG^-1 (SM tx + ru)
#pragma unroll
        for (int k = 0; k < NX; k++) st(zo(o_dx, i) + k, tx[k]);
#pragma unroll
        for (int k = 0; k < NU; k++) st(zo(o_dx, i) + NX + k, tu[k]);
        // theta(i+1) = (P tu + AM tx) + r2(i+1)
#pragma unroll
        for (int k = 0; k < NX; k++) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s1 = fma(ld(fb + oP + k + cc * NX), tu[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s2 = fma(ld(fb + oAM + k + cc * NX), tx[cc], s2);
          th[k] = (s1 + s2) + ld(o_ri + nz + (i + 1) * NX + k);
        }
      } else {
        // terminal stage :267-285
        double uc[NU], uN[NU], xN[NX];
        trsv_l<NU>(SG, ub, uc);
        trsv_lt<NU>(SG, uc, uN);
#pragma unroll
        for (int k = 0; k < NX; k++) {
          double sv = 0.0;
#pragma unroll
          for (int rr = 0; rr < NU; rr++) sv = fma(SM[rr][k], uN[rr], sv);
          sa[k] = tx[k] + sv;
        }
        trsv_lt<NX>(Mm, sa, sb);
#pragma unroll
        for (int k = 0; k < NX; k++) {
          xN[k] = -sb[k];
          sa[k] = xN[k] + th[k];
        }
        trsv_l<NX>(Lc, sa, sb);
        trsv_lt<NX>(Lc, sb, sc);
#pragma unroll
        for (int k = 0; k < NX; k++) {
          lp[k] = -sc[k];
          st(lo(o_dx, i) + k, lp[k]);
          st(zo(o_dx, i) + k, xN[k]);
        }
#pragma unroll
        for (int k = 0; k < NU; k++) st(zo(o_dx, i) + NX + k, uN[k]);
        double zN[NS];
#pragma unroll
        for (int k = 0; k < NX; k++) zN[k] = xN[k];
#pragma unroll
        for (int k = 0; k < NU; k++) zN[NX + k] = uN[k];
        finish_stage(i, zN);
      }
    }
    // backward recursion :297-327
    for (int i = N - 1; i >= 0; i--) {
      const int fb = o_fac + i * FS;
      double Lc[NX][NX], Mm[NX][NX], SG[NU][NU];
      load_lower(fb + oL, Lc);
      load_lower(fb + oM, Mm);
      load_lower_u(fb + oSG, SG);
      double ua[NU], ui[NU], sa[NX], sb[NX], sc[NX], xi_[NX];
#pragma unroll
      for (int k = 0; k < NU; k++) {
        double sv = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; rr++) sv = fma(ld(fb + oP + rr + k * NX), lp[rr], sv);
        ua[k] = ld(zo(o_dx, i) + NX + k) + sv;
      }
      trsv_lt<NU>(SG, ua, ui);
#pragma unroll
      for (int k = 0; k < NX; k++) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int rr = 0; rr < NU; rr++) s1 = fma(ld(fb + oSM + rr + k * NU), ui[rr], s1);
#pragma unroll
        for (int rr = 0; rr < NX; rr++) s2 = fma(ld(fb + oAM + rr + k * NX), lp[rr], s2);
        sa[k] = (ld(zo(o_dx, i) + k) + s1) + s2;
      }
      trsv_lt<NX>(Mm, sa, sb);
#pragma unroll
      for (int k = 0; k < NX; k++) {
        xi_[k] = -sb[k];
        sa[k] = ld(lo(o_dx, i) + k) + xi_[k];
      }
      trsv_l<NX>(Lc, sa, sb);
      trsv_lt<NX>(Lc, sb, sc);
      double zi[NS];
#pragma unroll
      for (int k = 0; k < NX; k++) {
        lp[k] = -sc[k];
        st(lo(o_dx, i) + k, lp[k]);
        st(zo(o_dx, i) + k, xi_[k]);
        zi[k] = xi_[k];
      }
#pragma unroll
      for (int k = 0; k < NU; k++) {
        st(zo(o_dx, i) + NX + k, ui[k]);
        zi[NX + k] = ui[k];
      }
      finish_stage(i, zi);
    }
  }
  // dv = (rv + gamma .* A dz) ./ mus ; dy = b - A dz for one stage (:331-341)
  __device__ __forceinline__ void finish_stage(int i, const double (&dz)[NS]) {
#pragma unroll
    for (int k = 0; k < NC; k++) {
      const double sv = Az_entry(dz, i, k);
      const double rv = ld(o_ri + nz + nl + i * NC + k);
      const double ga = ld(o_gm + i * NC + k), mu = ld(o_gm + nv + i * NC + k);
      st(vo(o_dx, i) + k, ((-rv) + ga * sv) / mu);
      st(yo(o_dx, i) + k, (-sv) + (-__ldg(d + i * NC + k)));
    }
  }

  // ProjectDuals on xi
  __device__ void project() {
    for (int i = 0; i <= N; i++)
#pragma unroll
      for (int k = 0; k < NC; k++) st(vo(o_xi, i) + k, fmax(ld(vo(o_xi, i) + k), 0.0));
  }

  // dx = xi - xk (y-aware), its norm, and FullFeasibility::CheckFeasibility on
  // it (full_feasibility.cc:25-88).  Returns the status in *feas.
  __device__ double diff_and_feasibility(double tol, bool check, int* feas) {
    double s[3] = {0, 0, 0};
    double mx0 = -INFINITY, mx1 = 0, mx2 = 0, mx3 = 0, mp0 = 0, mp1 = 0, mp2 = 0;
    double sm0 = 0, sm1 = 0;
    double zp[NS], lc[NX], ln[NX];
#pragma unroll
    for (int k = 0; k < NS; k++) zp[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NX; k++) {
      lc[k] = ld(lo(o_xi, 0) + k) + (-1.0) * ld(lo(o_xk, 0) + k);
      ln[k] = 0.0;
    }
    for (int i = 0; i <= N; i++) {
      double z[NS], v[NC];
#pragma unroll
      for (int k = 0; k < NS; k++) {
        z[k] = ld(zo(o_xi, i) + k) + (-1.0) * ld(zo(o_xk, i) + k);
        st(zo(o_dx, i) + k, z[k]);
        s[0] = fma(z[k], z[k], s[0]);
      }
#pragma unroll
      for (int k = 0; k < NX; k++) {
        st(lo(o_dx, i) + k, lc[k]);
        s[1] = fma(lc[k], lc[k], s[1]);
      }
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++)
          ln[k] = ld(lo(o_xi, i + 1) + k) + (-1.0) * ld(lo(o_xk, i + 1) + k);
      }
#pragma unroll
      for (int k = 0; k < NC; k++) {
        v[k] = ld(vo(o_xi, i) + k) + (-1.0) * ld(vo(o_xk, i) + k);
        st(vo(o_dx, i) + k, v[k]);
        s[2] = fma(v[k], v[k], s[2]);
        const double yv = ld(yo(o_xi, i) + k) + (-1.0) * ld(yo(o_xk, i) + k);
        st(yo(o_dx, i) + k, yv + (-__ldg(d + i * NC + k)));
      }
      if (check) {
        const double* Qm = Q + (size_t)i * NX * NX;
        const double* Rm = R + (size_t)i * NU * NU;
        const double* Sm = S + (size_t)i * NU * NX;
        const double* Em = E + (size_t)i * NC * NX;
        const double* Lm = L + (size_t)i * NC * NU;
#pragma unroll
        for (int k = 0; k < NC; k++) {
          mx0 = fmax(mx0, Az_entry(z, i, k));
          mp1 = fmax(mp1, fabs(v[k]));
          sm1 += (-__ldg(d + i * NC + k)) * v[k];
        }
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double gz, hh;
          if (i == 0) {
            gz = -z[rr];
            hh = -__ldg(x0 + rr);
          } else {
            const double* Ap = A + (size_t)(i - 1) * NX * NX;
            const double* Bp = B + (size_t)(i - 1) * NX * NU;
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) s1 = fma(__ldg(Ap + rr + cc * NX), zp[cc], s1);
#pragma unroll
            for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Bp + rr + cc * NX), zp[NX + cc], s2);
            gz = (s1 + s2) - z[rr];
            hh = -__ldg(c + (size_t)(i - 1) * NX + rr);
          }
          mx1 = fmax(mx1, fabs(gz));
          mp2 = fmax(mp2, fabs(lc[rr]));
          sm1 += hh * lc[rr];
        }
#pragma unroll
        for (int rr = 0; rr < NS; rr++) {
          double s1 = 0.0, s2 = 0.0, p, fe;
          if (rr < NX) {
#pragma unroll
            for (int cc = 0; cc < NX; cc++) s1 = fma(__ldg(Qm + rr + cc * NX), z[cc], s1);
#pragma unroll
            for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Sm + cc + rr * NU), z[NX + cc], s2);
            fe = __ldg(q + i * NX + rr);
            double sv = 0.0;
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(__ldg(Em + k + rr * NC), v[k], sv);
            p = sv + (-lc[rr]);
            if (i < N) {
              const double* Am = A + (size_t)i * NX * NX;
              double sa = 0.0;
#pragma unroll
              for (int cc = 0; cc < NX; cc++) sa = fma(__ldg(Am + cc + rr * NX), ln[cc], sa);
              p += sa;
            }
          } else {
            const int ru = rr - NX;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) s1 = fma(__ldg(Sm + ru + cc * NU), z[cc], s1);
#pragma unroll
            for (int cc = 0; cc < NU; cc++) s2 = fma(__ldg(Rm + ru + cc * NU), z[NX + cc], s2);
            fe = __ldg(r + i * NU + ru);
            double sv = 0.0;
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(__ldg(Lm + k + ru * NC), v[k], sv);
            p = sv;
            if (i < N) {
              const double* Bm = B + (size_t)i * NX * NU;
              double sa = 0.0;
#pragma unroll
              for (int cc = 0; cc < NX; cc++) sa = fma(__ldg(Bm + cc + ru * NX), ln[cc], sa);
              p += sa;
            }
          }
          mx2 = fmax(mx2, fabs(s1 + s2));
          mx3 = fmax(mx3, fabs(z[rr]));
          sm0 += fe * z[rr];
          mp0 = fmax(mp0, fabs(p));
        }
      }
#pragma unroll
      for (int k = 0; k < NS; k++) zp[k] = z[k];
#pragma unroll
      for (int k = 0; k < NX; k++) lc[k] = ln[k];
    }
    *feas = 0;
    if (check) {
      const double w = mx3;
      const bool dual_inf = (mx0 <= w * tol) && (mx1 <= tol * w) && (mx2 <= tol * w) &&
                            (sm0 < 0.0) && (w > 1e-14);
      const double u = fmax(mp1, mp2);
      const bool primal_inf = (mp0 <= tol * u) && (sm1 < 0.0);
      *feas = (primal_inf ? 1 : 0) + (dual_inf ? 2 : 0);
    }
    const double a = sqrt(s[0]), b = sqrt(s[1]), c2 = sqrt(s[2]);
    return sqrt(a * a + b * b + c2 * c2);
  }

  __device__ void copy_vars(int from, int to) {
    const int n = nz + nl + 2 * nv;
    for (int e = 0; e < n; e++) st(to + e, ld(from + e));
  }

  __device__ void write_result(int from, double* z, double* l, double* v, double* y) {
    for (int e = 0; e < nz; e++) z[e] = ld(from + e);
    for (int e = 0; e < nl; e++) l[e] = ld(from + nz + e);
    for (int e = 0; e < nv; e++) {
      v[e] = ld(from + nz + nl + e);
      y[e] = ld(from + nz + nl + nv + e);
    }
  }
};

template <int NX, int NU, int NC>
__global__ void __launch_bounds__(32, 8) mpc_lane_kernel(const __grid_constant__ LaneArgs a) {
  using LN = Lane<NX, NU, NC>;
  const fbstab_options& o = a.opts;
  const double sigma = o.sigma0, alpha = o.alpha;
  const int lane = threadIdx.x;
  LN p;
  p.N = a.N;
  p.nz = (a.N + 1) * LN::NS;
  p.nl = (a.N + 1) * NX;
  p.nv = (a.N + 1) * NC;
  const int VS = p.nz + p.nl + 2 * p.nv;
  p.ws = a.ws + (size_t)blockIdx.x * a.ws_stride + lane;
  p.o_xk = 0;
  p.o_xi = VS;
  p.o_dx = 2 * VS;
  p.o_ri = 3 * VS;
  p.o_gm = p.o_ri + p.nz + p.nl + p.nv;
  p.o_fac = p.o_gm + 2 * p.nv;

  // per-lane solver state (fbstab_algorithm-impl.h:113-304 as a phase machine)
  bool active = false, exhausted = false;
  int inst = 0, phase = PH_TOP;
  int eflag = FBSTAB_MAXITERATIONS, status = FBSTAB_STATUS_OK;
  int newton = 0, prox = 0, backtracks = 0, evals = 0;
  int k = 0, inner_i = 0, ls_j = 0;
  double E0 = 0, Ek = 0, last_rk = 0, inner_tol = 0, combo_tol = 0, dx_norm = 0;
  double merit[5] = {0, 0, 0, 0, 0};
  double Eo = 0, Ei_c = 0, Eo_c = 0, tstep = 1.0, m0 = 0, current_merit = 0;
  bool need_eval = false, pick_xi = false;

  for (;;) {
    // ---- idle lanes pull the next instance --------------------------------
    if (!active && !exhausted) {
      inst = atomicAdd(a.counter, 1);
      if (inst >= a.batch) {
        exhausted = true;
      } else {
        p.bind(a, inst);
        const double fn = p.init(a.z + (size_t)inst * p.nz, a.l + (size_t)inst * p.nl,
                                 a.v + (size_t)inst * p.nv);
        combo_tol = o.abs_tol + o.rel_tol * (1.0 + fn);
        dx_norm = sqrt((double)p.nz + (double)p.nl + (double)p.nv);  // dx_.Fill(1.0), impl:142
        eflag = FBSTAB_MAXITERATIONS;
        status = FBSTAB_STATUS_OK;
        newton = prox = backtracks = evals = 0;
        k = inner_i = ls_j = 0;
        E0 = Ek = last_rk = inner_tol = 0.0;
        phase = PH_TOP;
        need_eval = true;
        active = true;
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;

    // ---- evaluate -----------------------------------------------------------
    EvalOut e;
    e.Ei = e.Eo = 0.0;
    if (active && need_eval) {
      const bool self_bar = (phase == PH_TOP) || (phase == PH_FINAL);
      const int base = (phase == PH_FINAL && !pick_xi) ? p.o_xk : p.o_xi;
      e = p.evaluate(base, phase == PH_TRIAL, tstep, self_bar, sigma, alpha);
      evals++;
    }
    // ---- decide ---------------------------------------------------------------
    bool do_commit = false, do_newton = false, prox_end = false, finish = false;
    bool to_inner_top = false;
    int which = 0;  // 0: xk, 1: xi, 2: dx is the result
    if (active) {
      need_eval = false;
      if (phase == PH_TOP) {  // impl:158-185
        Ek = e.Eo;
        last_rk = Ek;
        bool bad = false;
        if (k == 0) {
          E0 = Ek;
          inner_tol = saturate(E0, o.inner_tol_min, o.inner_tol_max, &bad);
        }
        if (!bad && (Ek <= combo_tol || dx_norm <= o.stall_tol)) {
          eflag = FBSTAB_SUCCESS;
          finish = true;
        } else {
          if (!bad) inner_tol = saturate(inner_tol * o.delta, o.inner_tol_min, Ek, &bad);
          if (bad) {
            status = FBSTAB_STATUS_SATURATE;
            finish = true;
          } else {
#pragma unroll
            for (int m = 0; m < 5; m++) merit[m] = 0.0;
            Ei_c = e.Ei;
            Eo_c = e.Eo;
            inner_i = 0;
            to_inner_top = true;
          }
        }
      } else if (phase == PH_TRIAL) {  // Armijo test, impl:286-296
        const double mp = 0.5 * e.Ei * e.Ei;
        if (mp <= m0 - 2.0 * tstep * o.eta * current_merit) {
          do_commit = true;
          Ei_c = e.Ei;
          Eo_c = e.Eo;
          inner_i++;
          to_inner_top = true;
        } else {
          tstep *= o.beta;
          backtracks++;
          ls_j++;
          if (ls_j < o.max_linesearch_iters) {
            need_eval = true;  // next trial
          } else {
            // every trial failed: the step is still taken (impl:295-298)
            do_commit = true;
            inner_i++;
            if (inner_i < o.max_inner_iters) {
              phase = PH_REEVAL;
              need_eval = true;
            } else {
              prox_end = true;
            }
          }
        }
      } else if (phase == PH_REEVAL) {
        Ei_c = e.Ei;
        Eo_c = e.Eo;
        to_inner_top = true;
      } else {  // PH_FINAL
        last_rk = e.Eo;
        eflag = FBSTAB_MAXITERATIONS;
        which = pick_xi ? 1 : 0;
        finish = true;
      }
      if (to_inner_top) {  // top of an inner iteration, impl:237-260
        bool inner_done = (inner_i >= o.max_inner_iters);
        if (!inner_done) {
          Eo = Eo_c;
          last_rk = Eo;
          if ((Ei_c <= inner_tol && Eo < Ek) || (Ei_c <= o.inner_tol_min)) inner_done = true;
          if (newton >= o.max_newton_iters) inner_done = true;
        }
        if (inner_done)
          prox_end = true;
        else
          do_newton = true;
      }
    }
    // ---- commit the accepted (or forced) step: xi <- xi + t dx ---------------
    if (do_commit) p.commit(tstep);
    // ---- Newton step ------------------------------------------------------------
    if (do_newton) {
      if (!p.factor(sigma, alpha)) {  // impl:263-267
        status = FBSTAB_STATUS_FACTOR_FAILED;
        finish = true;
        which = 0;
      } else {
        p.solve();
        newton++;
        current_merit = 0.5 * Ei_c * Ei_c;
#pragma unroll
        for (int m = 4; m > 0; m--) merit[m] = merit[m - 1];
        merit[0] = current_merit;
        m0 = current_merit;
        if (o.nonmonotone_linesearch) {
#pragma unroll
          for (int m = 1; m < 5; m++) m0 = fmax(m0, merit[m]);
        }
        tstep = 1.0;
        ls_j = 0;
        phase = PH_TRIAL;
        need_eval = true;
      }
    }
    // ---- end of the subproblem, impl:300-216 -------------------------------------
    if (prox_end) {
      p.project();
      if (newton >= o.max_newton_iters) {  // impl:188-199
        pick_xi = Eo < Ek;
        phase = PH_FINAL;
        need_eval = true;
      } else {
        int feas = 0;
        dx_norm = p.diff_and_feasibility(o.infeas_tol, o.check_feasibility != 0, &feas);
        if (feas != 0) {
          eflag = (feas == 1)   ? FBSTAB_PRIMAL_INFEASIBLE
                  : (feas == 2) ? FBSTAB_DUAL_INFEASIBLE
                                : FBSTAB_PRIMAL_DUAL_INFEASIBLE;
          which = 2;
          finish = true;
        } else {
          p.copy_vars(p.o_xi, p.o_xk);
          prox++;
          k++;
          if (k >= o.max_prox_iters) {
            eflag = FBSTAB_MAXITERATIONS;
            which = 0;
            finish = true;
          } else {
            phase = PH_TOP;
            need_eval = true;
          }
        }
      }
    }
    // ---- WriteVariable + PrepareOutput, impl:349-383 ---------------------------
    if (finish) {
      const int from = which == 0 ? p.o_xk : which == 1 ? p.o_xi : p.o_dx;
      p.write_result(from, a.z + (size_t)inst * p.nz, a.l + (size_t)inst * p.nl,
                     a.v + (size_t)inst * p.nv, a.y + (size_t)inst * p.nv);
      fbstab_out* out = a.out + inst;
      out->eflag = eflag;
      out->newton_iters = newton;
      out->prox_iters = prox;
      out->status = status;
      out->residual = last_rk;
      out->initial_residual = E0;
      out->solve_time = -1.0;
      out->ls_backtracks = backtracks;
      out->residual_evals = evals;
      active = false;
    }
  }
}

typedef void (*LaneKernel)(const LaneArgs);
struct LaneVariant {
  int nx, nu, nc;
  LaneKernel fn;
};
const LaneVariant kLaneVariants[] = {
    {4, 1, 4, mpc_lane_kernel<4, 1, 4>},
    {2, 1, 6, mpc_lane_kernel<2, 1, 6>},
};

}  // namespace

bool MpcLaneSupported(int nx, int nu, int nc) {
  for (const LaneVariant& v : kLaneVariants)
    if (v.nx == nx && v.nu == nu && v.nc == nc) return true;
  return false;
}

size_t MpcLaneWsDoublesPerWarp(int N, int nx, int nu, int nc) {
  const size_t K = N + 1;
  const size_t nz = K * (nx + nu), nl = K * nx, nv = K * nc;
  const size_t fs = (size_t)nx * (nx + 1) + (size_t)nx * nx + 2 * (size_t)nx * nu +
                    (size_t)nu * (nu + 1) / 2;
  return 32 * (3 * (nz + nl + 2 * nv) + (nz + nl + nv) + 2 * nv + K * fs);
}

int MpcLaneLaunch(int N, int nx, int nu, int nc, int batch, int max_warps, const MpcData& data,
                  double* z, double* l, double* v, double* y, fbstab_out* out,
                  const fbstab_options& opts, double* ws, int* counter,
                  cudaStream_t stream) {
  LaneKernel fn = nullptr;
  for (const LaneVariant& var : kLaneVariants)
    if (var.nx == nx && var.nu == nu && var.nc == nc) fn = var.fn;
  if (!fn) return 1;
  LaneArgs a;
  a.N = N;
  a.batch = batch;
  a.data = data;
  a.z = z;
  a.l = l;
  a.v = v;
  a.y = y;
  a.out = out;
  a.ws = ws;
  a.ws_stride = MpcLaneWsDoublesPerWarp(N, nx, nu, nc);
  a.counter = counter;
  a.opts = opts;
  a.comp = -1;
  memset(&a.io, 0, sizeof(a.io));
  const int warps = std::min(max_warps, (batch + 31) / 32);
  fn<<<warps, 32, 0, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

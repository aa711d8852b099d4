// mpc_riccati.cuh -- Problem policy for MPC-structured QPs: one team (CTA) per
// instance walks the horizon; stage matrices and factor blocks move through
// shared-memory rings filled by the TMA engine (tma.cuh).
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
#include "mpc_riccati.h"
#include "tma.cuh"

// Diagonal slot of the warp-wide Cholesky: 0 = rsqrt(x) (rounds 1-2), 1 = 1 / sqrt(x) with
// both operations correctly rounded -- the reciprocal of the number the reference divides
// by; see FBSTAB_LANE_DIAG in mpc_lane.cu and profiles/r2_lane_diag_ab.txt: on the servo
// problem 94.9 % -> 97.9 % of the instances on the oracle's trajectory (its own FMA on/off
// floor: 97.6 %) for 4 % of throughput (cfg 4a40 18.8 k -> 18.0 k, cfg 4b 15.0 k -> 14.5 k).
#ifndef FBSTAB_CTA_DIAG
#define FBSTAB_CTA_DIAG 1
#endif

namespace fbs {

// ---- small dense helpers (column-major, ld = rows), team-cooperative -------

// Convention of every Cholesky factor in this file (and in mpc_lane.cu): the
// diagonal slot holds the RECIPROCAL of the pivot, so that the ~40 dependent
// divisions per stage of the triangular solves become multiplications (an FP64
// division is a chain of ~12 dependent operations).  Results differ from the
// dividing form by rounding only.
//
// Systems of at most 32 rows are factored and solved by ONE warp with
// warp-level synchronisation (the team's other warps wait at a single barrier
// afterwards): the team versions needed two CTA barriers per Cholesky column
// and one per substitution step, ~260 barriers per stage and Newton step at
// nx = 18, and the kernel spent most of its time in them
// (profiles/r1_mpc_cta_barriers.txt).

// In-place lower Cholesky (Eigen LLT unblocked order).  false on pivot <= 0.
__device__ __forceinline__ bool team_chol(const Team& t, int T, double* M, int m) {
  bool ok = true;
  if (m <= 32) {
    if (t.warp() == 0) {
      const int lane = t.lane();
      // lane i keeps the running diagonal M(i,i) - sum_{j<k} M(i,j)^2 in a
      // register, so the pivot chain of a step is one FMA, the reciprocal
      // square root, a shuffle and a multiply; the column's dot products do not
      // depend on the pivot and overlap with it
      double diag = (lane < m) ? M[lane + lane * m] : 1.0;
      for (int k = 0; k < m; k++) {
        double a = 0.0;
        if (lane > k && lane < m)
          for (int j = 0; j < k; j++) a = fma(M[lane + j * m], M[k + j * m], a);
        const double x = __shfl_sync(0xffffffffu, diag, k);
        if (!(x > 0.0)) ok = false;
#if FBSTAB_CTA_DIAG == 0
        const double rx = rsqrt(x);
#else
        const double rx = 1.0 / sqrt(x);  // as the lane kernel (FBSTAB_LANE_DIAG, mpc_lane.cu)
#endif
        if (lane > k && lane < m) {
          const double l = (M[lane + k * m] - a) * rx;
          M[lane + k * m] = l;
          diag = fma(-l, l, diag);
        }
        if (lane == k) M[k + k * m] = rx;
        __syncwarp();
      }
    }
    ok = __syncthreads_and(ok);
    return ok;
  }
  for (int k = 0; k < m; k++) {
    double s = 0.0;
    for (int j = 0; j < k; j++) {
      const double a = M[k + j * m];
      s = fma(a, a, s);
    }
    double x = M[k + k * m] - s;
    if (!(x > 0.0)) ok = false;
    const double rx = 1.0 / sqrt(x);
    for (int i = k + 1 + t.rank(); i < m; i += T) {
      double a = 0.0;
      for (int j = 0; j < k; j++) a = fma(M[i + j * m], M[k + j * m], a);
      M[i + k * m] = (M[i + k * m] - a) * rx;
    }
    t.sync();
    if (t.rank() == 0) M[k + k * m] = rx;  // after everyone has read the old pivot
  }
  t.sync();
  return ok;
}

// y <- L^-1 x (lower, reciprocal diagonal, column oriented).  x is destroyed.
__device__ __forceinline__ void trsv_l(const Team& t, int T, const double* L, int m, double* x,
                              double* y) {
  if (m <= 32) {
    if (t.warp() == 0) {
      const int lane = t.lane();
      double xi = (lane < m) ? x[lane] : 0.0;
      for (int j = 0; j < m; j++) {
        const double xj = __shfl_sync(0xffffffffu, xi, j) * L[j + j * m];
        if (lane == j) xi = xj;
        if (lane > j && lane < m) xi = fma(-L[lane + j * m], xj, xi);
      }
      if (lane < m) y[lane] = xi;
    }
    t.sync();
    return;
  }
  for (int j = 0; j < m; j++) {
    const double xj = x[j] * L[j + j * m];
    if (t.rank() == 0) y[j] = xj;
    for (int i = j + 1 + t.rank(); i < m; i += T)
      x[i] = fma(-L[i + j * m], xj, x[i]);
    t.sync();
  }
}
// y <- L^-T x.  x is destroyed.
__device__ __forceinline__ void trsv_lt(const Team& t, int T, const double* L, int m, double* x,
                               double* y) {
  if (m <= 32) {
    if (t.warp() == 0) {
      const int lane = t.lane();
      double xr = (lane < m) ? x[lane] : 0.0;
      for (int i = m - 1; i >= 0; i--) {
        const double xi = __shfl_sync(0xffffffffu, xr, i) * L[i + i * m];
        if (lane == i) xr = xi;
        if (lane < i) xr = fma(-L[i + lane * m], xi, xr);
      }
      if (lane < m) y[lane] = xr;
    }
    t.sync();
    return;
  }
  for (int i = m - 1; i >= 0; i--) {
    const double xi = x[i] * L[i + i * m];
    if (t.rank() == 0) y[i] = xi;
    for (int r = t.rank(); r < i; r += T)
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
    X[r + j * rows] = s * L[j + j * m];
  }
}

// The same with the row kept in registers (compile-time m): the shared-memory
// version re-reads every X(r,k) it has just stored, a store -> load round trip
// per term of a strictly sequential recurrence.
template <int M>
__device__ __forceinline__ void row_trsm_lt_reg(const double* L, const double* src, double* X,
                                                int rows, int r) {
  double xr[M];
#pragma unroll
  for (int j = 0; j < M; j++) {
    double s = src[r + j * rows];
#pragma unroll
    for (int k = 0; k < j; k++) s = fma(-xr[k], L[j + k * M], s);
    xr[j] = s * L[j + j * M];
  }
#pragma unroll
  for (int j = 0; j < M; j++) X[r + j * rows] = xr[j];
}

constexpr int kRing = 3;  // slots per ring (prefetch distance 2 stages)
#ifndef FBS_MPC_DATA_RING
#define FBS_MPC_DATA_RING 2
#endif
constexpr int kDataRing = FBS_MPC_DATA_RING;  // stage-data ring, prefetch distance kDataRing-1: one stage
                                              // (20-40 k cycles) covers the copy; 2 slots measured 4 % faster
                                              // than 3 on the copolymerisation shape (8.7 KB less per CTA)

// Views of one stage's matrices (shared-memory ring slot or global memory).
struct StageData {
  const double *Q, *R, *S, *A, *B, *E, *L;
};

// KNX, KNU, KNC, KT: compile-time stage sizes and team size (0 = run-time value).
// The four OCP shapes of the BASELINE configs are instantiated with constants,
// which unrolls every stage loop and removes the index arithmetic -- the
// generic code measured 7.4k warp instructions per stage at nx=4
// (profiles/r1_mpc_generic_instruction_bound.txt).
#define FBS_MPC_DIMS                                   \
  const int nx = KNX ? KNX : this->nx;                 \
  const int nu = KNU ? KNU : this->nu;                 \
  const int nc = KNC ? KNC : this->nc;                 \
  (void)nx; (void)nu; (void)nc;
#define FBS_MPC_TEAM const int T = KT ? KT : t.size(); (void)T;

template <int KNX, int KNU, int KNC, int KT>
struct MpcProblem {
  int N, nx, nu, nc, nz, nl, nv;
  MpcLayout lay;
  // this instance's sequences
  const double *Q, *R, *S, *q, *r, *A, *B, *c, *E, *L, *d, *x0;
  // per-CTA workspace
  double *gamma, *mus;             // nv each
  double *Gam, *tv;                // aliases of dx.y / dx.v (dead while they are used)
  double *Qt, *Rt, *St, *Linv;     // stage temporaries (shared)
  double *sa, *sb, *sc;            // max(nx,nu) each (shared)
  double* fac;                     // (N+1) factor blocks: shared (resident) or global
  double* fslot0;                  // factor ring: kRing slots of FS doubles (shared)
  double* dslot0;                  // stage-data ring: kRing slots of SD doubles (shared)
  unsigned bar0;                   // shared address of the 2*kRing mbarriers (data, factor)
  unsigned fphase, dphase;         // per-slot wait parity bits (persist across calls)

  // block / slot offsets (same arithmetic as Footprint() in mpc_riccati.cu;
  // compile-time constants in the specialised instantiations)
  struct FacOff { int L, M, AM, SM, P, SG, FS; };
  struct DataOff { int Q, R, S, A, B, E, L, SD; };
  __device__ __forceinline__ FacOff fac_off() const {
    FBS_MPC_DIMS
    const int nxx = nx * nx, nux = nu * nx;
    FacOff o;
    o.L = 0; o.M = nxx; o.AM = 2 * nxx; o.SM = 3 * nxx; o.P = 3 * nxx + nux;
    o.SG = 3 * nxx + 2 * nux;
    o.FS = (3 * nxx + 2 * nux + nu * nu + 1) & ~1;
    return o;
  }
  __device__ __forceinline__ DataOff data_off() const {
    FBS_MPC_DIMS
    auto sub = [](int n) { return (n + 2) & ~1; };
    DataOff o;
    o.Q = 0;
    o.R = o.Q + sub(nx * nx);
    o.S = o.R + sub(nu * nu);
    o.A = o.S + sub(nu * nx);
    o.B = o.A + sub(nx * nx);
    o.E = o.B + sub(nu * nx);
    o.L = o.E + sub(nc * nx);
    o.SD = o.L + sub(nc * nu);
    return o;
  }
  __device__ __forceinline__ double* fslot(int s) const { return fslot0 + (size_t)s * fac_off().FS; }
  __device__ __forceinline__ double* dslot(int s) const { return dslot0 + (size_t)s * data_off().SD; }
  __device__ __forceinline__ unsigned dbar(int s) const { return bar0 + 8u * (unsigned)s; }
  __device__ __forceinline__ unsigned fbar(int s) const { return bar0 + 8u * (unsigned)(kRing + s); }
  __device__ __forceinline__ int ns() const { FBS_MPC_DIMS return nx + nu; }
  __device__ __forceinline__ double b(int i) const { return -d[i]; }
  // entries of the stacked f = [q(i); r(i)] and h = -[x0; c(0); ...; c(N-1)]
  __device__ __forceinline__ double fvec(int e) const {
    const int nsv = ns(), i = e / nsv;
    return f_entry(i, e - i * nsv);
  }
  __device__ __forceinline__ double hvec(int k) const {
    FBS_MPC_DIMS
    const int i = k / nx;
    return h_entry(i, k - i * nx);
  }
  __device__ __forceinline__ const double* Qi(int i) const { FBS_MPC_DIMS return Q + (size_t)i * nx * nx; }
  __device__ __forceinline__ const double* Ri(int i) const { FBS_MPC_DIMS return R + (size_t)i * nu * nu; }
  __device__ __forceinline__ const double* Si(int i) const { FBS_MPC_DIMS return S + (size_t)i * nu * nx; }
  __device__ __forceinline__ const double* Ai(int i) const { FBS_MPC_DIMS return A + (size_t)i * nx * nx; }
  __device__ __forceinline__ const double* Bi(int i) const { FBS_MPC_DIMS return B + (size_t)i * nx * nu; }
  __device__ __forceinline__ const double* Ei(int i) const { FBS_MPC_DIMS return E + (size_t)i * nc * nx; }
  __device__ __forceinline__ const double* Lci(int i) const { FBS_MPC_DIMS return L + (size_t)i * nc * nu; }

  // mpc_data.h:89-97
  __device__ double forcing_norm(const Team& t) const {
    FBS_MPC_DIMS
    FBS_MPC_TEAM
    double s[1] = {0.0};
    for (int i = t.rank(); i < (N + 1) * nx; i += T) s[0] += q[i] * q[i];
    for (int i = t.rank(); i < (N + 1) * nu; i += T) s[0] += r[i] * r[i];
    for (int i = t.rank(); i < (N + 1) * nc; i += T) s[0] += d[i] * d[i];
    for (int i = t.rank(); i < nx; i += T) s[0] += x0[i] * x0[i];
    for (int i = t.rank(); i < N * nx; i += T) s[0] += c[i] * c[i];
    team_sum(t, s);
    return sqrt(s[0]);
  }

  // (E(i) x(i) + L(i) u(i))[k]   -- one entry of A_qp z, mpc_data.cc:66-105
  __device__ __forceinline__ double Az_entry(const double* z, int i, int k) const {
    FBS_MPC_DIMS
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
    FBS_MPC_DIMS
    FBS_MPC_TEAM
    for (int e = t.rank(); e < nv; e += T) {
      const int i = e / nc, k = e - i * nc;
      y[e] = -d[e] - Az_entry(z, i, k);
    }
    t.sync();
  }

  // (H z)[idx], mpc_data.cc:17-64
  __device__ __forceinline__ double Hz_entry(const double* z, int i, int rr) const {
    FBS_MPC_DIMS
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
    FBS_MPC_DIMS
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
    FBS_MPC_DIMS
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
    FBS_MPC_DIMS
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
    FBS_MPC_DIMS
    return rr < nx ? q[(size_t)i * nx + rr] : r[(size_t)i * nu + rr - nx];
  }
  __device__ __forceinline__ double h_entry(int i, int rr) const {
    FBS_MPC_DIMS
    return i == 0 ? -x0[rr] : -c[(size_t)(i - 1) * nx + rr];
  }

  // tz = ((f + Hz) + G'l) + A'v ; tl = h - Gz
  __device__ __noinline__ void kkt(const Team& t, const Vars& x, double* oz, double* ol) const {
    FBS_MPC_DIMS
    FBS_MPC_TEAM
    const int nsv = ns();
    for (int e = t.rank(); e < nz + nl; e += T) {
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

  // ---- TMA rings ---------------------------------------------------------------
  // Stage-data ring: thread 0 issues the copies of stage i's seven matrices
  // into slot i % kRing two stages ahead of their use; everyone waits on the
  // slot's mbarrier before reading.
  __device__ __forceinline__ void issue_stage_data(int i) {
    FBS_MPC_DIMS
    const int s = i % kDataRing;
    double* base = dslot(s);
    const unsigned bar = dbar(s);
    const int nxx = nx * nx, nuu = nu * nu, nux = nu * nx, ncx = nc * nx, ncu = nc * nu;
    const double* src[7] = {Qi(i), Ri(i), Si(i), Ai(i), Bi(i), Ei(i), Lci(i)};
    const DataOff dof = data_off();
    const int off[7] = {dof.Q, dof.R, dof.S, dof.A, dof.B, dof.E, dof.L};
    const int cnt[7] = {nxx, nuu, nux, i < N ? nxx : 0, i < N ? nux : 0, ncx, ncu};
    unsigned bytes = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) bytes += tma::copy_run(base + off[k], src[k], cnt[k], bar, false);
    tma::mbar_arrive_expect_tx(bar, bytes);
#pragma unroll
    for (int k = 0; k < 7; k++) tma::copy_run(base + off[k], src[k], cnt[k], bar, true);
  }
  __device__ __forceinline__ StageData stage_data(int i) {
    StageData sd;
    if (!lay.data_ring) {
      sd.Q = Qi(i); sd.R = Ri(i); sd.S = Si(i); sd.A = Ai(i); sd.B = Bi(i);
      sd.E = Ei(i); sd.L = Lci(i);
      return sd;
    }
    const int s = i % kDataRing;
    tma::mbar_wait(dbar(s), (dphase >> s) & 1u);
    dphase ^= 1u << s;
    const double* base = dslot(s);
    // a run the TMA engine could not take whole is read in place (tma.cuh)
    auto at = [&](const double* src, int off, int n) {
      return tma::bulk_able(src, n) ? base + off : src;
    };
    const DataOff dof = data_off();
    const int nxx = nx * nx, nuu = nu * nu, nux = nu * nx, ncx = nc * nx, ncu = nc * nu;
    sd.Q = at(Qi(i), dof.Q, nxx); sd.R = at(Ri(i), dof.R, nuu); sd.S = at(Si(i), dof.S, nux);
    sd.A = at(Ai(i), dof.A, nxx); sd.B = at(Bi(i), dof.B, nux); sd.E = at(Ei(i), dof.E, ncx);
    sd.L = at(Lci(i), dof.L, ncu);
    return sd;
  }
  // Factor ring (streaming mode): block i lives in slot i % kRing.
  __device__ __forceinline__ void issue_factor_block(int i) {
    const int s = i % kRing;
    const unsigned bytes = (unsigned)fac_off().FS * 8u;
    tma::mbar_arrive_expect_tx(fbar(s), bytes);
    tma::bulk_g2s(tma::smem_addr(fslot(s)), fac + (size_t)i * fac_off().FS, bytes, fbar(s));
  }
  __device__ __forceinline__ void wait_factor_block(int i) {
    const int s = i % kRing;
    tma::mbar_wait(fbar(s), (fphase >> s) & 1u);
    fphase ^= 1u << s;
  }
  // where block i is built (factor) / read (solve)
  __device__ __forceinline__ double* block(int i) const {
    return lay.fac_smem ? fac + (size_t)i * fac_off().FS : fslot(i % kRing);
  }

  // RiccatiLinearSolver::Initialize, riccati_linear_solver.cc:77-210
  __device__ __noinline__ bool factor(const Team& t, const Vars& x, const Vars& xbar,
                         double sigma, double alpha) {
    FBS_MPC_DIMS
    FBS_MPC_TEAM
    for (int i = t.rank(); i < nv; i += T) {
      const double ys = x.y[i] + sigma * (x.v[i] - xbar.v[i]);
      double ga, mu;
      pfb_barrier(ys, x.v[i], alpha, sigma, &ga, &mu);
      gamma[i] = ga;
      mus[i] = mu;
      Gam[i] = div_nr(ga, mu);
    }
    const int nxx = nx * nx, nuu = nu * nu, nux = nu * nx;
    const FacOff fo = fac_off();
    const bool stream = !lay.fac_smem;
    if (lay.data_ring && t.rank() == 0) {
      for (int j = 0; j < kDataRing - 1 && j <= N; j++) issue_stage_data(j);
    }
    // L(0) = sqrt(sigma) I, :127
    {
      const double rs = 1.0 / sqrt(sigma);  // reciprocal-diagonal convention
      double* L0 = block(0) + fo.L;
      for (int e = t.rank(); e < nxx; e += T) L0[e] = (e % nx == e / nx) ? rs : 0.0;
    }
    t.sync();
    bool ok = true;
    for (int i = 0; i <= N; i++) {
      double* Bk = block(i);
      double* Li = Bk + fo.L;
      double* Mi = Bk + fo.M;
      double* AMi = Bk + fo.AM;
      double* SMi = Bk + fo.SM;
      double* Pi = Bk + fo.P;
      double* SGi = Bk + fo.SG;
      if (lay.data_ring && t.rank() == 0 && i + kDataRing - 1 <= N)
        issue_stage_data(i + kDataRing - 1);
      const StageData sd = stage_data(i);
      const double* Gi = Gam + (size_t)i * nc;
      // barrier-augmented stage Hessian (:102-123) and Linv = inv(L L') (:142-144)
      const int work = nxx + nuu + nux + nx;
      for (int e = t.rank(); e < work; e += T) {
        if (e < nxx) {
          const int rr = e % nx, cc = e / nx;
          if (rr >= cc) {
            double s = 0.0;
            for (int k = 0; k < nc; k++)
              s = fma(sd.E[k + rr * nc], Gi[k] * sd.E[k + cc * nc], s);
            Qt[e] = (sd.Q[e] + (rr == cc ? sigma : 0.0)) + s;
          }
        } else if (e < nxx + nuu) {
          const int f = e - nxx;
          const int rr = f % nu, cc = f / nu;
          if (rr >= cc) {
            double s = 0.0;
            for (int k = 0; k < nc; k++)
              s = fma(sd.L[k + rr * nc], Gi[k] * sd.L[k + cc * nc], s);
            Rt[f] = (sd.R[f] + (rr == cc ? sigma : 0.0)) + s;
          }
        } else if (e < nxx + nuu + nux) {
          const int f = e - nxx - nuu;
          const int rr = f % nu, cc = f / nu;
          double s = 0.0;
          for (int k = 0; k < nc; k++)
            s = fma(sd.L[k + rr * nc], Gi[k] * sd.E[k + cc * nc], s);
          St[f] = sd.S[f] + s;
        } else {
          // column cc of inv(L L'): forward then backward substitution
          const int cc = e - nxx - nuu - nux;
          if (KNX > 0) {  // compile-time size: the column lives in registers
            constexpr int NXR = KNX > 0 ? KNX : 1;
            double wr[NXR];
#pragma unroll
            for (int k = 0; k < NXR; k++) wr[k] = (k == cc) ? 1.0 : 0.0;
#pragma unroll
            for (int j = 0; j < NXR; j++) {
              wr[j] *= Li[j + j * NXR];
              const double wj = wr[j];
#pragma unroll
              for (int k = j + 1; k < NXR; k++) wr[k] = fma(-Li[k + j * NXR], wj, wr[k]);
            }
#pragma unroll
            for (int k = NXR - 1; k >= 0; k--) {
              double s = wr[k];
#pragma unroll
              for (int j = k + 1; j < NXR; j++) s = fma(-Li[j + k * NXR], wr[j], s);
              wr[k] = s * Li[k + k * NXR];
            }
#pragma unroll
            for (int k = 0; k < NXR; k++) Linv[(size_t)cc * NXR + k] = wr[k];
            continue;
          }
          double* w = Linv + (size_t)cc * nx;
          for (int k = 0; k < nx; k++) w[k] = (k == cc) ? 1.0 : 0.0;
          for (int j = 0; j < nx; j++) {
            w[j] *= Li[j + j * nx];
            const double wj = w[j];
            for (int k = j + 1; k < nx; k++) w[k] = fma(-Li[k + j * nx], wj, w[k]);
          }
          for (int k = nx - 1; k >= 0; k--) {
            double s = w[k];
            for (int j = k + 1; j < nx; j++) s = fma(-Li[j + k * nx], w[j], s);
            w[k] = s * Li[k + k * nx];
          }
        }
      }
      t.sync();
      // M = chol(Qt + Linv), :145-147
      for (int e = t.rank(); e < nxx; e += T) {
        const int rr = e % nx, cc = e / nx;
        Mi[e] = (rr >= cc) ? Qt[e] + Linv[e] : 0.0;
      }
      t.sync();
      ok = team_chol(t, T, Mi, nx) && ok;
      // AM = A M^-T, SM = St M^-T, :149-161
      for (int w = t.rank(); w < nx + nu; w += T) {
        if (KNX > 0) {
          constexpr int NXR = KNX > 0 ? KNX : 1;
          if (w < nx) {
            if (i < N) row_trsm_lt_reg<NXR>(Mi, sd.A, AMi, nx, w);
          } else {
            row_trsm_lt_reg<NXR>(Mi, St, SMi, nu, w - nx);
          }
        } else if (w < nx) {
          if (i < N) row_trsm_lt(Mi, nx, sd.A, AMi, nx, w);
        } else {
          row_trsm_lt(Mi, nx, St, SMi, nu, w - nx);
        }
      }
      t.sync();
      // SG = chol(Rt - SM SM'), :163-166
      for (int e = t.rank(); e < nuu; e += T) {
        const int rr = e % nu, cc = e / nu;
        double s = 0.0;
        if (rr >= cc) {
          for (int k = 0; k < nx; k++) s = fma(SMi[rr + k * nu], SMi[cc + k * nu], s);
          s = Rt[e] - s;
        }
        SGi[e] = s;
      }
      t.sync();
      ok = team_chol(t, T, SGi, nu) && ok;
      if (i < N) {
        // P = (AM SM' - B) SG^-T, :170-175
        for (int rr = t.rank(); rr < nx; rr += T) {
          for (int j = 0; j < nu; j++) {
            double s = 0.0;
            for (int k = 0; k < nx; k++) s = fma(AMi[rr + k * nx], SMi[j + k * nu], s);
            Pi[rr + j * nx] = s - sd.B[rr + j * nx];
          }
          if (KNU > 0) {
            constexpr int NUR = KNU > 0 ? KNU : 1;
            row_trsm_lt_reg<NUR>(SGi, Pi, Pi, nx, rr);
          } else {
            row_trsm_lt(SGi, nu, Pi, Pi, nx, rr);
          }
        }
        // the slot that receives L(i+1) may still be the source of the bulk
        // store of block i-2
        if (stream && t.rank() == 0) tma::bulk_wait_read_all();
        t.sync();
        // L(i+1) = chol(sigma I + P P' + AM AM'), :179-183
        double* Ln = block(i + 1) + fo.L;
        for (int e = t.rank(); e < nxx; e += T) {
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
        ok = team_chol(t, T, Ln, nx) && ok;
      }
      if (stream) {
        // block i is complete: shared -> global behind the sweep (TMA store)
        tma::fence_proxy_async();
        t.sync();
        if (t.rank() == 0) {
          tma::bulk_s2g(fac + (size_t)i * fo.FS, tma::smem_addr(Bk), (unsigned)fo.FS * 8u);
          tma::bulk_commit();
        }
      }
    }
    if (stream) {
      if (t.rank() == 0) tma::bulk_wait_all();
      t.sync();
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

  // RiccatiLinearSolver::Solve on r = -(rz,rl,rv), riccati_linear_solver.cc:212-344.
  // Storage reuse (every alias is dead while it is borrowed):
  //   r1 = r.z - A' r3 overwrites rz in place; r2 is rl itself;
  //   r3 = r.v ./ mus lives in dx.v; theta(i) in dx.l(i); M^-1 h(i) in dx.z's
  //   x(i) slot and SG^-1(SM tx + ru) in its u(i) slot until the backward sweep
  //   replaces them with the step.
  __device__ __noinline__ void solve(const Team& t, double* rz, const double* rl,
                        const double* rv, const Vars& dx) {
    FBS_MPC_DIMS
    FBS_MPC_TEAM
    const int nsv = ns();
    const FacOff fo = fac_off();
    const bool stream = !lay.fac_smem;
    if (stream && t.rank() == 0) {
      issue_factor_block(0);
      if (N >= 1) issue_factor_block(1);
    }
    for (int i = t.rank(); i < nv; i += T) tv[i] = div_nr(-rv[i], mus[i]);
    t.sync();
    for (int e = t.rank(); e < nz; e += T) {
      const int i = e / nsv, rr = e - i * nsv;
      rz[e] = (-rz[e]) - ATv_entry(tv, i, rr);
    }
    for (int k = t.rank(); k < nx; k += T) dx.l[k] = rl[k];  // theta(0) = r2(0)
    t.sync();
    const double* r1 = rz;
    // forward recursion :232-285
    for (int i = 0; i <= N; i++) {
      if (stream) {
        if (t.rank() == 0 && i + 2 <= N) issue_factor_block(i + 2);
        wait_factor_block(i);
      }
      const double* Bk = block(i);
      const double* Li = Bk + fo.L;
      const double* Mi = Bk + fo.M;
      const double* AMi = Bk + fo.AM;
      const double* SMi = Bk + fo.SM;
      const double* Pi = Bk + fo.P;
      const double* SGi = Bk + fo.SG;
      double* tx = dx.z + (size_t)i * nsv;
      double* tu = tx + nx;
      double* th = dx.l + (size_t)i * nx;
      // h(i) = (L L')^-1 theta(i) - r1x(i)
      for (int k = t.rank(); k < nx; k += T) sa[k] = th[k];
      t.sync();
      trsv_l(t, T, Li, nx, sa, sb);
      trsv_lt(t, T, Li, nx, sb, sc);
      for (int k = t.rank(); k < nx; k += T) sa[k] = sc[k] - r1[(size_t)i * nsv + k];
      t.sync();
      trsv_l(t, T, Mi, nx, sa, tx);  // tx = M^-1 h
      for (int k = t.rank(); k < nu; k += T)
        sb[k] = mv_row(SMi, nu, nx, tx, k) + r1[(size_t)i * nsv + nx + k];
      t.sync();
      if (i < N) {
        trsv_l(t, T, SGi, nu, sb, tu);  // tu = SG^-1 (SM tx + ru)
        for (int k = t.rank(); k < nx; k += T)
          dx.l[(size_t)(i + 1) * nx + k] =
              (mv_row(Pi, nx, nu, tu, k) + mv_row(AMi, nx, nx, tx, k)) +
              rl[(size_t)(i + 1) * nx + k];
        t.sync();
      } else {
        // terminal stage :267-285
        double* uN = tu;
        double* xN = tx;
        double* lN = th;
        trsv_l(t, T, SGi, nu, sb, sc);
        trsv_lt(t, T, SGi, nu, sc, uN);
        for (int k = t.rank(); k < nx; k += T)
          sa[k] = tx[k] + mtv_col(SMi, nu, uN, k);
        t.sync();
        trsv_lt(t, T, Mi, nx, sa, sb);
        for (int k = t.rank(); k < nx; k += T) {
          const double xv = -sb[k];
          sa[k] = xv + th[k];
          xN[k] = xv;
        }
        t.sync();
        trsv_l(t, T, Li, nx, sa, sb);
        trsv_lt(t, T, Li, nx, sb, sc);
        for (int k = t.rank(); k < nx; k += T) lN[k] = -sc[k];
        t.sync();
      }
    }
    // backward recursion :297-327 (tx and tu were kept from the forward pass;
    // the reference recomputes the same values).  Blocks N-1 and N-2 are still
    // in their ring slots.
    for (int i = N - 1; i >= 0; i--) {
      if (stream) {
        // slot (i-2) % kRing holds block i+1, consumed one stage ago
        if (t.rank() == 0 && i - 2 >= 0) issue_factor_block(i - 2);
        if (i <= N - 3) wait_factor_block(i);
      }
      const double* Bk = block(i);
      const double* Li = Bk + fo.L;
      const double* Mi = Bk + fo.M;
      const double* AMi = Bk + fo.AM;
      const double* SMi = Bk + fo.SM;
      const double* Pi = Bk + fo.P;
      const double* SGi = Bk + fo.SG;
      const double* lp = dx.l + (size_t)(i + 1) * nx;
      double* xi = dx.z + (size_t)i * nsv;
      double* ui = xi + nx;
      double* li = dx.l + (size_t)i * nx;
      for (int k = t.rank(); k < nu; k += T)
        sa[k] = ui[k] + mtv_col(Pi, nx, lp, k);
      t.sync();
      trsv_lt(t, T, SGi, nu, sa, ui);
      for (int k = t.rank(); k < nx; k += T)
        sa[k] = (xi[k] + mtv_col(SMi, nu, ui, k)) + mtv_col(AMi, nx, lp, k);
      t.sync();
      trsv_lt(t, T, Mi, nx, sa, sb);
      for (int k = t.rank(); k < nx; k += T) {
        const double xv = -sb[k];
        sa[k] = li[k] + xv;
        xi[k] = xv;
      }
      t.sync();
      trsv_l(t, T, Li, nx, sa, sb);
      trsv_lt(t, T, Li, nx, sb, sc);
      for (int k = t.rank(); k < nx; k += T) li[k] = -sc[k];
      t.sync();
    }
    // dv = (rv + gamma .* A dz) ./ mus ; dy = b - A dz   (:331-341)
    for (int e = t.rank(); e < nv; e += T) {
      const int i = e / nc, k = e - i * nc;
      const double s = Az_entry(dx.z, i, k);
      dx.v[e] = div_nr((-rv[e]) + gamma[e] * s, mus[e]);
      dx.y[e] = (-s) + (-d[e]);
    }
    t.sync();
  }

  // FullFeasibility::CheckFeasibility, full_feasibility.cc:25-88
  __device__ __noinline__ int feasibility(const Team& t, const Vars& dx, double tol) {
    FBS_MPC_DIMS
    FBS_MPC_TEAM
    const int nsv = ns();
    double mx[4] = {-INFINITY, 0.0, 0.0, 0.0};
    double mp[3] = {0.0, 0.0, 0.0};
    double sm[2] = {0.0, 0.0};
    for (int e = t.rank(); e < nv; e += T) {
      const int i = e / nc, k = e - i * nc;
      mx[0] = fmax(mx[0], Az_entry(dx.z, i, k));
      mp[1] = fmax(mp[1], fabs(dx.v[e]));
      sm[1] += (-d[e]) * dx.v[e];
    }
    for (int e = t.rank(); e < nl; e += T) {
      const int i = e / nx, rr = e - i * nx;
      mx[1] = fmax(mx[1], fabs(Gz_entry(dx.z, i, rr)));
      mp[2] = fmax(mp[2], fabs(dx.l[e]));
      sm[1] += h_entry(i, rr) * dx.l[e];
    }
    for (int e = t.rank(); e < nz; e += T) {
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

// dense_small2.cu -- TWO warps per problem for small dense QPs (nz <= 32,
// nl <= 8, nv <= 64): the round-2 successor of the warp-per-problem kernel in
// dense_small.cu.
//
// Why: the warp kernel is latency bound with nothing to switch to.  Shared
// memory (27.8 KB of A, H, G per instance) caps an SM at EIGHT instances, i.e.
// two warps per scheduler; ncu (profiles/r2_dense_small_v3_ncu.txt) shows issue
// slots 38% busy, the FP64 pipe 35%, and every step of the 32-step elimination
// waiting on a store -> barrier -> load -> reciprocal chain.  The instances per
// SM cannot grow, so the warps per INSTANCE do: the same eight slabs, sixteen
// warps, and every phase of the solve split between the two warps of a team:
//
//  * constraint-indexed vectors (v, y, Gamma, mu, rv): thread t of the team owns
//    entry t (64 threads, 64 constraints); z- and l-type vectors are replicated
//    in both warps (lane j owns entry j in each);
//  * mat-vecs: each warp sums over its half of the contraction index (A'v over
//    its 32 rows of A, H z over its 16 columns) and the two partial sums meet in
//    shared memory; A z is one row per thread;
//  * E = H + sigma I + A' Gamma A: each warp accumulates the ten lower 8x8 blocks
//    over ITS 32 rows of A on the FP64 tensor cores (80 DMMAs each instead of
//    160), the partials are added block row by block row on the way to the row
//    layout of the elimination;
//  * Gauss-Jordan elimination of [E | G' | rhs]: lane i of both warps owns row i,
//    warp h owns the 8-column blocks h and h+2 of E and half of the augmented
//    columns, so the pivot-row broadcast of a step -- the dominant cost of the
//    warp kernel -- is split in two; the pivot COLUMN travels through shared
//    memory one step ahead, next to the pivot row.
// A team synchronises with a named barrier (bar.sync id, 64).  The algorithm is
// the one of dense_small.cu / engine.cuh (fbstab_algorithm-impl.h:113-304,
// dense_cholesky_solver.cc:32-148, full_residual.cc:49-118,
// full_feasibility.cc:25-88); the sums are split differently, so results agree
// with the warp kernel to rounding, not bit for bit.
#include <cstring>

#include "common.cuh"
#include "dense_small.h"
#include "engine.cuh"

namespace fbs {
namespace small2 {

constexpr int NZ = 32, NL = 8, NV = 64, LD = 34;
constexpr int kTeams = 8;  // instances in flight per CTA (two warps each)

// slab of one team (doubles): same data layout as the warp kernel
constexpr int OFF_H = 0;  // packed lower triangle Hp[i(i+1)/2 + j]
constexpr int H_SIZE = NZ * (NZ + 1) / 2;
constexpr int OFF_A = OFF_H + H_SIZE;    // As[j + LD*k] = A(k,j)
constexpr int OFF_G = OFF_A + NV * LD;   // Gs[j + LD*r] = G(r,j)
constexpr int OFF_ZB = OFF_G + NL * LD;  // broadcast copies z(32) l(8) v(64)
constexpr int OFF_LB = OFF_ZB + NZ;
constexpr int OFF_VB = OFF_LB + NL;
constexpr int OFF_SCR = OFF_VB + NV;
constexpr int SCR_SIZE = 392;
constexpr int SLAB = OFF_SCR + SCR_SIZE;
// scratch map: [0, 320) phase scratch (Gamma | block-row transposition T | pivot
// buffers | Y), [320, 392) exchange area: two 32-vectors would not fit next to the
// scalars, so the vector exchange uses [320, 384) and the scalar slots [384, 392)
constexpr int EXV = 320;  // 2 x 32 partial vectors
constexpr int EXS = 384;  // 8 scalars
// pivot buffers of the elimination (per step parity): half-0 part (16 + 6),
// half-1 part (16 + 4), pivot column (32)
constexpr int PB_H1 = 22, PB_COL = 44, PB_STRIDE = 76;
static_assert(2 * PB_STRIDE <= EXV, "pivot buffers must not reach the exchange area");
static_assert((size_t)SLAB * kTeams * sizeof(double) <= 232448, "slabs must fit one SM");
static_assert(SLAB % 2 == 0 && OFF_A % 2 == 0 && OFF_G % 2 == 0 && OFF_ZB % 2 == 0 &&
                  OFF_VB % 2 == 0 && OFF_SCR % 2 == 0 && PB_H1 % 2 == 0 && PB_COL % 2 == 0 &&
                  PB_STRIDE % 2 == 0,
              "16-byte accesses need 16-byte aligned targets");

struct Args {
  int nz, nl, nv, batch;
  const double *H, *f, *G, *h, *A, *b;
  double *z, *l, *v, *y;
  fbstab_out* out;
  int* counter;
  fbstab_options opts;
  int comp;
  fbstab_component_io io;
};

struct V {  // iterate: z, l replicated in both warps; v, y one entry per thread
  double z, l, v, y;
};
struct R {
  double z, l, v;
};

// ---- shared-space accessors -----------------------------------------------------
__device__ __forceinline__ double lds(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts(unsigned a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts2(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts_if(bool p, unsigned a, double x) {
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}" ::"r"(a),
               "d"(x), "r"((int)p)
               : "memory");
}
__device__ __forceinline__ void sts2_if(bool p, unsigned a, double x, double y) {
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.s32 q, %3, 0;\n @q st.shared.v2.f64 [%0], {%1,%2};\n}" ::"r"(a),
      "d"(x), "d"(y), "r"((int)p)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ constexpr unsigned D(int i) { return 8u * (unsigned)i; }
__device__ __forceinline__ double div_r(double a, double b, double rb) {
  const double q = a * rb;
  const double rem = fma(-b, q, a);
  return fma(rem, rb, q);
}
__device__ __forceinline__ void pfb_barrier_r(double ys, double v, double alpha, double sigma,
                                              double* gamma, double* mu) {
  const double r = sqrt(ys * ys + v * v);
  double ga, gb;
  if (r < 1e-13) {
    const double d = 0.70710678118654752440;
    ga = alpha * (1.0 - d);
    gb = ga;
  } else {
    const double rr = 1.0 / r;
    const double qa = div_r(ys, r, rr), qb = div_r(v, r, rr);
    if (ys > 0.0 && v > 0.0) {
      ga = alpha * (1.0 - qa) + (1.0 - alpha) * v;
      gb = alpha * (1.0 - qb) + (1.0 - alpha) * ys;
    } else {
      ga = alpha * (1.0 - qa);
      gb = alpha * (1.0 - qb);
    }
  }
  *gamma = ga;
  *mu = gb + sigma * ga;
}

struct Team {
  unsigned sb;  // shared byte address of the team's slab
  int lane, half, tid, bar;
  int nz, nl, nv;
  double fr, hr, br;  // f(lane), h(lane), b(tid)
  double gamma, mus, rmu;
  double a[16];  // row `lane` of E: own column blocks (half, half + 2), rotating
  double g[6];   // own augmented entries: half 0 G'(:,0..3) and the rhs, half 1 G'(:,4..7)
  double dinv;
  bool ok;

  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ unsigned scr(int i) const { return sb + D(OFF_SCR + i); }

  // z, l (replicated) and v (one entry per thread) into the broadcast area
  __device__ __forceinline__ void publish(const V& x) {
    sync();
    if (half == 0) {
      sts(sb + D(OFF_ZB + lane), x.z);
      if (lane < NL) sts(sb + D(OFF_LB + lane), x.l);
    }
    sts(sb + D(OFF_VB + tid), x.v);
    sync();
  }
  __device__ __forceinline__ double Hel(int i, int j) const {
    const int r = max(i, j), c = min(i, j);
    return lds(sb + D(OFF_H + (r * (r + 1) >> 1) + c));
  }
  // this warp's share of (A' vb)[lane] (rows 32 half .. 32 half + 31 of A) and, with
  // WITH_H, of (H zb)[lane] (columns 16 half .. 16 half + 15): 8 trips, 6 chains
  template <bool WITH_H>
  __device__ __forceinline__ double part_ATv_Hz(double* hz_part) const {
    const unsigned ac = sb + D(OFF_A + LD * 32 * half + lane);
    const unsigned vb = sb + D(OFF_VB + 32 * half);
    const unsigned zb = sb + D(OFF_ZB);
    const unsigned hrow = sb + D(OFF_H + (lane * (lane + 1) >> 1));
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll 2
    for (int t = 0; t < 8; t++) {
      const int k = 4 * t, j = 16 * half + 2 * t;
      const double2 v0 = lds2(vb + D(k));
      const double2 v1 = lds2(vb + D(k + 2));
      const double a0 = lds(ac + D(LD * k)), a1 = lds(ac + D(LD * (k + 1)));
      const double a2 = lds(ac + D(LD * (k + 2))), a3 = lds(ac + D(LD * (k + 3)));
      if (WITH_H) {
        const double2 zz = lds2(zb + D(j));
        const double e0 = lds(j <= lane ? hrow + D(j) : sb + D(OFF_H + (j * (j + 1) >> 1) + lane));
        const double e1 = lds(j + 1 <= lane ? hrow + D(j + 1)
                                            : sb + D(OFF_H + ((j + 1) * (j + 2) >> 1) + lane));
        h0 = fma(e0, zz.x, h0);
        h1 = fma(e1, zz.y, h1);
      }
      s0 = fma(a0, v0.x, s0);
      s1 = fma(a1, v0.y, s1);
      s2 = fma(a2, v1.x, s2);
      s3 = fma(a3, v1.y, s3);
    }
    if (WITH_H) *hz_part = h0 + h1;
    return (s0 + s1) + (s2 + s3);
  }
  // (G' lb)[lane]
  __device__ __forceinline__ double GTl() const {
    const unsigned gc = sb + D(OFF_G + lane);
    const unsigned lb = sb + D(OFF_LB);
    double s0 = 0.0;
#pragma unroll
    for (int r = 0; r < NL; r += 2) {
      const double2 ll = lds2(lb + D(r));
      s0 = fma(lds(gc + D(LD * r)), ll.x, s0);
      s0 = fma(lds(gc + D(LD * (r + 1))), ll.y, s0);
    }
    return s0;
  }
  // (G zb)[lane % 8], same value in the four lanes sharing lane % 8 (both warps)
  __device__ __forceinline__ double Gz() const {
    const int r = lane & 7, q = lane >> 3;
    const unsigned gr = sb + D(OFF_G + LD * r + 8 * q);
    const unsigned zb = sb + D(OFF_ZB + 8 * q);
    double s0 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 gg = lds2(gr + D(j));
      const double2 zz = lds2(zb + D(j));
      s0 = fma(gg.x, zz.x, s0);
      s0 = fma(gg.y, zz.y, s0);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    return s0;
  }
  // (A zb)[tid]: one row per thread
  __device__ __forceinline__ double Az1() const {
    const unsigned zb = sb + D(OFF_ZB);
    const unsigned arow = sb + D(OFF_A + LD * tid);
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll 4
    for (int j = 0; j < NZ; j += 4) {
      const double2 z0 = lds2(zb + D(j)), z1 = lds2(zb + D(j + 2));
      const double2 a0 = lds2(arow + D(j)), a1 = lds2(arow + D(j + 2));
      t0 = fma(a0.x, z0.x, t0);
      t1 = fma(a0.y, z0.y, t1);
      t2 = fma(a1.x, z1.x, t2);
      t3 = fma(a1.y, z1.y, t3);
    }
    return (t0 + t1) + (t2 + t3);
  }
  // both warps receive p0 + p1 of their per-lane partials (evaluated identically)
  __device__ __forceinline__ double combine(double part) {
    sts(scr(EXV + 32 * half + lane), part);
    sync();
    const double p0 = lds(scr(EXV + lane)), p1 = lds(scr(EXV + 32 + lane));
    return p0 + p1;
  }
  // both warps receive the two per-warp scalar pairs (slot s: 0 or 4)
  __device__ __forceinline__ void combine2(int s, double x, double y, double2* t0, double2* t1) {
    if (lane == 0) sts2(scr(EXS + s + 2 * half), x, y);
    sync();
    *t0 = lds2(scr(EXS + s));
    *t1 = lds2(scr(EXS + s + 2));
  }

  // ---- fused residual evaluation (see engine.cuh) ---------------------------------
  __device__ __forceinline__ EvalOut evaluate(const V& x, const V& xbar, double sigma,
                                              double alpha, R* ri) {
    publish(x);
    double hzp;
    double part = part_ATv_Hz<true>(&hzp) + hzp;
    if (half == 1) part += GTl();
    const double tz = fr + combine(part);
    const double gz = Gz();
    const double tl = (lane < NL) ? hr - gz : 0.0;
    const double rz = tz + sigma * (x.z - xbar.z);
    const double rl = tl + sigma * (x.l - xbar.l);
    ri->z = rz;
    ri->l = rl;
    double si = 0.0, so = 0.0;
    if (half == 0) {  // the replicated parts are counted once
      si = fma(rl, rl, rz * rz);
      so = fma(tl, tl, tz * tz);
    }
    const double ys = x.y + sigma * (x.v - xbar.v);
    const double r = pfb(ys, x.v, alpha);
    ri->v = r;
    si = fma(r, r, si);
    const double n = pnr(x.y, x.v, alpha);
    so = fma(n, n, so);
    double2 t0, t1;
    combine2(0, warp_sum(si), warp_sum(so), &t0, &t1);
    EvalOut e;
    e.Ei = sqrt(t0.x + t1.x);
    e.Eo = sqrt(t0.y + t1.y);
    return e;
  }

  // ---- Gauss-Jordan elimination, column split ---------------------------------------
  // One step.  NE: live entries of this warp's part of the rows (for the owner of
  // the pivot column they start with the pivot column itself and ROTate left by
  // one); NA augmented entries.  The pivot row and the raw pivot column of the NEXT
  // step are stored while the update produces them (predicated stores between the
  // DFMAs), so a step costs one team barrier.
  template <int NE, bool ROT, int NA>
  __device__ __forceinline__ void gj_step(int k, unsigned cb, unsigned nb) {
    constexpr int NP = NE / 2, NAP = (NA + 1) / 2;
    const int owner = (k >> 3) & 1;
    const unsigned mine = cb + D(half ? PB_H1 : 0), nmine = nb + D(half ? PB_H1 : 0);
    sync();
    double2 c[NP > 0 ? NP : 1], xr[NAP];
#pragma unroll
    for (int p = 0; p < NP; p++) c[p] = lds2(mine + D(2 * p));
#pragma unroll
    for (int r = 0; r < NAP; r++) xr[r] = lds2(mine + D(16 + 2 * r));
    // pivot and this lane's entry of the pivot column
    double d, aik;
    if (ROT) {
      d = c[0].x;
      aik = a[0];
    } else {
      d = lds(cb + D(owner ? PB_H1 : 0));
      aik = lds(cb + D(PB_COL + lane));
    }
    if (!(fabs(d) > 0.0)) ok = false;
    const double rd = 1.0 / d;
    if (lane == k) dinv = rd;
    const double lik = (lane != k) ? aik * rd : 0.0;
    const bool nxt = (lane == k + 1);
    if (ROT) {
#pragma unroll
      for (int m = 1; m < NE; m++) {
        const double cm = (m & 1) ? c[m / 2].y : c[m / 2].x;
        a[m - 1] = fma(-lik, cm, a[m]);
        if (m & 1) {
          if (m >= 3) sts2_if(nxt, nmine + D(m - 3), a[m - 3], a[m - 2]);
        }
      }
      a[NE - 1] = 0.0;
      sts2_if(nxt, nmine + D(NE - 2), a[NE - 2], 0.0);
    } else {
#pragma unroll
      for (int m = 0; m < NE; m++) {
        const double cm = (m & 1) ? c[m / 2].y : c[m / 2].x;
        a[m] = fma(-lik, cm, a[m]);
        if (m & 1) sts2_if(nxt, nmine + D(m - 1), a[m - 1], a[m]);
      }
    }
#pragma unroll
    for (int r = 0; r < NA; r++) {
      const double xm = (r & 1) ? xr[r / 2].y : xr[r / 2].x;
      g[r] = fma(-lik, xm, g[r]);
      if (r & 1) sts2_if(nxt, nmine + D(16 + r - 1), g[r - 1], g[r]);
    }
    if (NA & 1) sts2_if(nxt, nmine + D(16 + NA - 1), g[NA - 1], 0.0);
    // the raw pivot column of step k + 1 comes from the warp that owns column k + 1:
    // its (possibly just rotated) first live entry
    if (NE > 0) sts_if((((k + 1) >> 3) & 1) == half && k + 1 < NZ, nb + D(PB_COL + lane), a[0]);
  }
  template <int NE0, bool ROT0, int NE1, bool ROT1>
  __device__ __forceinline__ void gj_segment(int k0) {
#pragma unroll 1
    for (int k = k0; k < k0 + 8; k++) {
      const unsigned cb = scr(PB_STRIDE * (k & 1)), nb = scr(PB_STRIDE * ((k + 1) & 1));
      if (half == 0)
        gj_step<NE0, ROT0, 5>(k, cb, nb);
      else
        gj_step<NE1, ROT1, 4>(k, cb, nb);
    }
  }

  // ---- Newton step: LinearSolver::Initialize + ::Solve fused -----------------------
  __device__ __forceinline__ bool newton_step(const V& x, const V& xbar, double sigma,
                                              double alpha, const R& ri, V* dx) {
    const int r8 = lane >> 2, c4 = lane & 3;
    ok = true;
    {
      const double ys = x.y + sigma * (x.v - xbar.v);
      pfb_barrier_r(ys, x.v, alpha, sigma, &gamma, &mus);
      rmu = 1.0 / mus;
      const double Gam = div_r(gamma, mus, rmu);
      const double r2 = div_r(-ri.v, mus, rmu);
      sync();
      sts(sb + D(OFF_VB + tid), r2);
      sts(scr(tid), Gam);  // Gamma(k), k = tid: read back by this warp only
      sync();
    }
    // r1z = -rz - A'(rv / mus)
    const double rhs = (-ri.z) - combine(part_ATv_Hz<false>(nullptr));

    // E, lower 8x8 blocks, over this warp's 32 rows of A.  The four rows of a DMMA
    // chunk are 4 apart (k0, k0+4, k0+8, k0+12): the fragment loads of the four c4
    // groups then fall into two bank groups instead of colliding (LD = 34).
    {
      double C[4][4][2];
#pragma unroll
      for (int I = 0; I < 4; I++)
#pragma unroll
        for (int J = 0; J <= I; J++) {
          const int hr_ = 8 * I + r8, hc = 8 * J + 2 * c4;
          if (half == 0) {
            C[I][J][0] = Hel(hr_, hc) + ((I == J && r8 == 2 * c4) ? sigma : 0.0);
            C[I][J][1] = Hel(hr_, hc + 1) + ((I == J && r8 == 2 * c4 + 1) ? sigma : 0.0);
          } else {
            C[I][J][0] = 0.0;
            C[I][J][1] = 0.0;
          }
        }
      // chunk q = 0..7: rows k = 32 half + 16 (q >> 2) + (q & 3) + 4 c4
      const int kb = 32 * half + 4 * c4;
      const unsigned ar0 = sb + D(OFF_A + LD * kb + r8);
      const unsigned gk0 = scr(kb);
      double an[4], gn;
      gn = lds(gk0);
#pragma unroll
      for (int X = 0; X < 4; X++) an[X] = lds(ar0 + D(8 * X));
#pragma unroll 1
      for (int q = 0; q < 8; q++) {
        double af[4], bf[4];
#pragma unroll
        for (int X = 0; X < 4; X++) {
          af[X] = an[X];
          bf[X] = gn * an[X];
        }
        const int qn = (q + 1) & 7;
        const int kn = 16 * (qn >> 2) + (qn & 3);
        gn = lds(gk0 + D(kn));
#pragma unroll
        for (int X = 0; X < 4; X++) an[X] = lds(ar0 + D(LD * kn + 8 * X));
#pragma unroll
        for (int I = 0; I < 4; I++)
#pragma unroll
          for (int J = 0; J <= I; J++) dmma(C[I][J][0], C[I][J][1], af[I], bf[J]);
      }
      // block row by block row: warp 1's partial joins warp 0's in T (row layout, the
      // scratch behind Gamma), then every lane picks its own columns of its row --
      // blocks on and left of the diagonal from the row, right of it from the column
      // (E is symmetric)
      sync();  // both warps are past their Gamma reads
#pragma unroll
      for (int I = 0; I < 4; I++) {
        if (half == 1) {
#pragma unroll
          for (int J = 0; J <= I; J++)
            sts2(scr(LD * r8 + 8 * J + 2 * c4), C[I][J][0], C[I][J][1]);
        }
        sync();
        if (half == 0) {
#pragma unroll
          for (int J = 0; J <= I; J++) {
            const double2 t = lds2(scr(LD * r8 + 8 * J + 2 * c4));
            sts2(scr(LD * r8 + 8 * J + 2 * c4), C[I][J][0] + t.x, C[I][J][1] + t.y);
          }
        }
        sync();
        if ((lane >> 3) == I) {
          const unsigned row = scr(LD * (lane & 7));
#pragma unroll
          for (int bq = 0; bq < 2; bq++) {
            const int J = half + 2 * bq;  // own column block
            if (J <= I) {
#pragma unroll
              for (int jj = 0; jj < 8; jj += 2) {
                const double2 t = lds2(row + D(8 * J + jj));
                a[8 * bq + jj] = t.x;
                a[8 * bq + jj + 1] = t.y;
              }
            }
          }
        } else if ((lane >> 3) < I) {
          // E(lane, 8 I + rr) = E(8 I + rr, lane): column `lane` of the block row
          if (I == half) {
#pragma unroll
            for (int rr = 0; rr < 8; rr++) a[rr] = lds(scr(LD * rr + lane));
          } else if (I == half + 2) {
#pragma unroll
            for (int rr = 0; rr < 8; rr++) a[8 + rr] = lds(scr(LD * rr + lane));
          }
        }
        sync();
      }
    }
    // augmented columns: G' (lane j owns column j of G) and the right-hand side
#pragma unroll
    for (int r = 0; r < 4; r++) g[r] = lds(sb + D(OFF_G + LD * (4 * half + r) + lane));
    g[4] = rhs;  // (warp 1 carries it along unused: NA = 4 there)
    g[5] = 0.0;

    // pivot row 0 and pivot column 0 into buffer 0 (the loop's first barrier orders them)
    {
      const unsigned b0 = scr(0), mine = b0 + D(half ? PB_H1 : 0);
      const bool own = (lane == 0);
#pragma unroll
      for (int m = 0; m < 16; m += 2) sts2_if(own, mine + D(m), a[m], a[m + 1]);
      sts2_if(own, mine + D(16), g[0], g[1]);
      sts2_if(own, mine + D(18), g[2], g[3]);
      if (half == 0) {
        sts2_if(own, mine + D(20), g[4], 0.0);
        sts(b0 + D(PB_COL + lane), a[0]);
      }
    }
    gj_segment<16, true, 16, false>(0);
    gj_segment<8, false, 16, true>(8);
    gj_segment<8, true, 8, false>(16);
    gj_segment<0, false, 8, true>(24);
    // lane i now holds d_i * (E^-1 [G' a])(i, own columns) in g[] and dinv = 1 / d_i

    // Schur complement S = -sigma I - G Y, rhs c - G t with [Y t] = E^-1 [G' a]
    sync();  // the last step's buffer reads are done: Ys may overwrite them
    if (half == 0) {
      sts2(scr(10 * lane), g[0] * dinv, g[1] * dinv);
      sts2(scr(10 * lane + 2), g[2] * dinv, g[3] * dinv);
      sts2(scr(10 * lane + 8), g[4] * dinv, 0.0);
    } else {
      sts2(scr(10 * lane + 4), g[0] * dinv, g[1] * dinv);
      sts2(scr(10 * lane + 6), g[2] * dinv, g[3] * dinv);
    }
    sync();
    double dl = 0.0;
    if (half == 0) {
      double S0 = 0.0, S1 = 0.0, T0 = 0.0, T1 = 0.0;
#pragma unroll 2
      for (int kc = 0; kc < NZ / 4; kc++) {
        const int i = 4 * kc + c4;
        const double ge = lds(sb + D(OFF_G + LD * r8 + i));
        const double ye = lds(scr(10 * i + r8));
        const double te = lds(scr(10 * i + NL));
        dmma(S0, S1, ge, ye);
        dmma(T0, T1, ge, te);
      }
      __syncwarp();
      sts2(scr(EXV + 8 * r8 + 2 * c4), S0, S1);
      if (c4 == 0) sts(scr(EXS + r8), T0);
      __syncwarp();
      // 8 x 8 Gauss-Jordan by shuffles in lanes 0..7 (replicated in the other lanes)
      const int r = lane & 7;
      double srow[NL];
#pragma unroll
      for (int j = 0; j < NL; j += 2) {
        const double2 t = lds2(scr(EXV + 8 * r + j));
        srow[j] = -t.x - ((j == r) ? sigma : 0.0);
        srow[j + 1] = -t.y - ((j + 1 == r) ? sigma : 0.0);
      }
      double rhs2 = ((lane < NL) ? ri.l : 0.0) - lds(scr(EXS + r));
      double dsi = 0.0;
#pragma unroll 1
      for (int k = 0; k < NL; k++) {
        const double dk = __shfl_sync(0xffffffffu, srow[0], k);
        if (!(fabs(dk) > 0.0)) ok = false;
        const double rdk = 1.0 / dk;
        if (r == k) dsi = rdk;
        const double lrk = (r != k) ? srow[0] * rdk : 0.0;
#pragma unroll
        for (int j = 1; j < NL; j++) {
          const double cj = __shfl_sync(0xffffffffu, srow[j], k);
          srow[j - 1] = fma(-lrk, cj, srow[j]);
        }
        srow[NL - 1] = 0.0;
        const double rk = __shfl_sync(0xffffffffu, rhs2, k);
        rhs2 = fma(-lrk, rk, rhs2);
      }
      dl = (lane < NL) ? rhs2 * dsi : 0.0;
      if (lane < NL) sts(sb + D(OFF_LB + lane), dl);
      // a zero pivot here must reach warp 1 too
      if (lane == 0) sts(scr(EXS), ok ? 1.0 : 0.0);
    }
    sync();
    if (lds(scr(EXS)) == 0.0) ok = false;
    // dz = (t - Y dl): each warp subtracts its four columns, then the parts meet
    double acc = (half == 0) ? g[4] : 0.0;
    {
      const double2 d0 = lds2(sb + D(OFF_LB + 4 * half)), d1 = lds2(sb + D(OFF_LB + 4 * half + 2));
      acc = fma(-g[0], d0.x, acc);
      acc = fma(-g[1], d0.y, acc);
      acc = fma(-g[2], d1.x, acc);
      acc = fma(-g[3], d1.y, acc);
    }
    sync();  // EXS flag and S are consumed before the exchange area is reused
    const double dz = combine(acc) * dinv;
    dx->z = dz;
    dx->l = (lane < NL) ? lds(sb + D(OFF_LB + lane)) : 0.0;
    // dv = (rv + gamma .* (A dz)) ./ mus ; dy = b - A dz
    if (half == 0) sts(sb + D(OFF_ZB + lane), dz);
    sync();
    const double adz = Az1();
    dx->v = div_r(gamma * adz + (-ri.v), mus, rmu);
    dx->y = br - adz;
    // both warps must agree on failure
    double2 t0, t1;
    combine2(4, ok ? 0.0 : 1.0, 0.0, &t0, &t1);
    return (t0.x + t1.x) == 0.0;
  }

  // FullFeasibility::CheckFeasibility on dx
  __device__ __forceinline__ int feasibility(const V& dx, double tol) {
    publish(dx);
    const double adz = Az1();
    double d1 = (tid < nv) ? adz : -INFINITY;
    const double gz = Gz();
    const double d2 = (lane < NL) ? fabs(gz) : 0.0;
    double hzp;
    double atv = part_ATv_Hz<true>(&hzp);
    if (half == 1) atv += GTl();
    // two per-lane partial pairs: the phase scratch is free outside the Newton step
    sts(scr(32 * half + lane), hzp);
    sts(scr(64 + 32 * half + lane), atv);
    sync();
    const double hz = lds(scr(lane)) + lds(scr(32 + lane));
    const double atg = lds(scr(64 + lane)) + lds(scr(96 + lane));
    const double d3 = fabs(hz);
    const double w = warp_max(fabs(dx.z));
    const double d4 = warp_sum(fr * dx.z);
    const double p1 = warp_max(fabs(atg));
    // constraint-indexed parts: per-warp partials, then combined
    double p2 = warp_sum(br * dx.v);
    double um = warp_max(fabs(dx.v));
    d1 = warp_max(d1);
    if (lane == 0) {
      sts2(scr(EXS + 2 * half), p2, um);
      sts(scr(EXS + 4 + half), d1);
    }
    sync();
    const double2 q0 = lds2(scr(EXS)), q1 = lds2(scr(EXS + 2));
    const double2 dd = lds2(scr(EXS + 4));
    p2 = (q0.x + q1.x) + warp_sum((lane < NL) ? hr * dx.l : 0.0);
    const double umax = fmax(fmax(q0.y, q1.y), warp_max(fabs(dx.l)));
    const double D1 = fmax(dd.x, dd.y), D2 = warp_max(d2), D3 = warp_max(d3);
    const bool dual_inf = (D1 <= w * tol) && (D2 <= tol * w) && (D3 <= tol * w) && (d4 < 0.0) &&
                          (w > 1e-14);
    const bool primal_inf = (p1 <= tol * umax) && (p2 < 0.0);
    sync();  // scratch and scalar slots are free again
    return (primal_inf ? 1 : 0) + (dual_inf ? 2 : 0);
  }

  // ---- data staging: the two warps take alternate batches of 256 elements ------------
  __device__ __forceinline__ void stage_matrix(const double* src, int rows, int cols, int off) {
    const int n = rows * cols;
    for (int e0 = 256 * half; e0 < n; e0 += 512) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + lane + 32 * u;
        t[u] = (e < n) ? __ldg(src + e) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + lane + 32 * u;
        if (e < n) {
          const int c = e / rows, r = e - c * rows;
          sts(sb + D(off + LD * r + c), t[u]);
        }
      }
    }
  }
  __device__ __forceinline__ void stage_lower(const double* src, int rows) {
    const int n = rows * rows;
    for (int e0 = 256 * half; e0 < n; e0 += 512) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + lane + 32 * u;
        t[u] = (e < n) ? __ldg(src + e) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + lane + 32 * u;
        if (e < n) {
          const int c = e / rows, r = e - c * rows;
          if (r >= c) sts(sb + D(OFF_H + (r * (r + 1) >> 1) + c), t[u]);
        }
      }
    }
  }
  __device__ __forceinline__ void load_exact(const Args& a_, int inst) {
    const double* Asrc = a_.A + (size_t)inst * NV * NZ;
    const double* Hsrc = a_.H + (size_t)inst * NZ * NZ;
    const double* Gsrc = a_.G + (size_t)inst * NL * NZ;
    // A(r, c) at r + 64 c -> As[c + LD r]; e = lane + 32 u: c = u / 2, r = lane + 32 (u & 1)
#pragma unroll 1
    for (int u0 = 32 * half; u0 < 32 * half + 32; u0 += 16) {
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; u++) t[u] = __ldg(Asrc + lane + 32 * (u0 + u));
      const unsigned base = sb + D(OFF_A + LD * lane + (u0 >> 1));
#pragma unroll
      for (int u = 0; u < 16; u++) sts(base + D(LD * 32 * (u & 1) + (u >> 1)), t[u]);
    }
    // lower triangle of H: e = lane + 32 u: c = u, r = lane
    {
      const int u0 = 16 * half;
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; u++) t[u] = (lane >= u0 + u) ? __ldg(Hsrc + lane + 32 * (u0 + u)) : 0.0;
      const unsigned base = sb + D(OFF_H + (lane * (lane + 1) >> 1) + u0);
#pragma unroll
      for (int u = 0; u < 16; u++)
        if (lane >= u0 + u) sts(base + D(u), t[u]);
    }
    // G(r, c) at r + 8 c: e = lane + 32 u: c = lane / 8 + 4 u, r = lane % 8
    {
      double t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) t[u] = __ldg(Gsrc + lane + 32 * (4 * half + u));
      const unsigned base = sb + D(OFF_G + LD * (lane & 7) + (lane >> 3) + 16 * half);
#pragma unroll
      for (int u = 0; u < 4; u++) sts(base + D(4 * u), t[u]);
    }
  }
  __device__ __forceinline__ void load(const Args& a_, int inst) {
    sync();
    if (nz == NZ && nl == NL && nv == NV) {
      load_exact(a_, inst);
    } else {
      for (int e = 2 * tid; e < OFF_ZB; e += 128) sts2(sb + D(e), 0.0, 0.0);
      sync();
      stage_lower(a_.H + (size_t)inst * nz * nz, nz);
      stage_matrix(a_.A + (size_t)inst * nv * nz, nv, nz, OFF_A);
      stage_matrix(a_.G + (size_t)inst * nl * nz, nl, nz, OFF_G);
    }
    fr = (lane < nz) ? __ldg(a_.f + (size_t)inst * nz + lane) : 0.0;
    hr = (lane < nl) ? __ldg(a_.h + (size_t)inst * nl + lane) : 0.0;
    br = (tid < nv) ? __ldg(a_.b + (size_t)inst * nv + tid) : 0.0;
    sync();
  }
  __device__ __forceinline__ void prefetch(const Args& a_, int inst) const {
    const char* H = (const char*)(a_.H + (size_t)inst * nz * nz);
    const char* A = (const char*)(a_.A + (size_t)inst * nv * nz);
    const char* G = (const char*)(a_.G + (size_t)inst * nl * nz);
    for (int o = 128 * tid; o < 8 * nz * nz; o += 128 * 64) prefetch_l2(H + o);
    for (int o = 128 * tid; o < 8 * nv * nz; o += 128 * 64) prefetch_l2(A + o);
    for (int o = 128 * tid; o < 8 * nl * nz; o += 128 * 64) prefetch_l2(G + o);
  }
  // next instance index: thread 0 of the team pulls it, both warps read it
  __device__ __forceinline__ int next_instance(int* counter) {
    sync();
    if (tid == 0) sts(scr(EXS + 7), (double)atomicAdd(counter, 1));
    sync();
    return (int)lds(scr(EXS + 7));
  }
};

__device__ __forceinline__ void v_axpy(const Team& w, const V& src, double a, const V& dx,
                                       V* dst) {
  dst->z = src.z + a * dx.z;
  dst->l = src.l + a * dx.l;
  dst->v = src.v + a * dx.v;
  const double y = src.y + a * dx.y;
  dst->y = y + (-a) * w.br;
}
__device__ __forceinline__ V v_select(bool c, const V& a, const V& b) {
  V r;
  r.z = c ? a.z : b.z;
  r.l = c ? a.l : b.l;
  r.v = c ? a.v : b.v;
  r.y = c ? a.y : b.y;
  return r;
}

enum Phase { P_TOP = 0, P_TRIAL = 1, P_REEVAL = 2, P_FINAL = 3 };

// FBstabAlgorithm::Solve for one instance, executed by the two warps of a team in
// lockstep: every decision is taken on values both warps hold bit for bit.
__device__ __forceinline__ void solve_one(Team& w, const Args& A, int inst) {
  const fbstab_options& o = A.opts;
  const int lane = w.lane, tid = w.tid;
  const double sigma = o.sigma0, alpha = o.alpha;
  V xk, xi, xp, dx, xe;
  R ri;
  xk.z = (lane < w.nz) ? A.z[(size_t)inst * w.nz + lane] : 0.0;
  xk.l = (lane < w.nl) ? A.l[(size_t)inst * w.nl + lane] : 0.0;
  xk.v = (tid < w.nv) ? A.v[(size_t)inst * w.nv + tid] : 0.0;
  // forcing norm and margin y = b - A z
  double combo_tol;
  {
    double fn = (w.half == 0) ? fma(w.hr, w.hr, w.fr * w.fr) : 0.0;  // hr = 0 beyond nl
    fn = fma(w.br, w.br, fn);
    double2 t0, t1;
    w.sync();
    w.combine2(0, warp_sum(fn), 0.0, &t0, &t1);
    combo_tol = o.abs_tol + o.rel_tol * (1.0 + sqrt(t0.x + t1.x));
  }
  xk.y = 0.0;
  w.publish(xk);
  xk.y = w.br - w.Az1();
  xi = xk;
  xp = xk;
  dx = xk;
  double dx_norm = sqrt((double)w.nz + (double)w.nl + (double)w.nv);
  int eflag = FBSTAB_MAXITERATIONS, status = FBSTAB_STATUS_OK;
  int newton = 0, prox = 0, backtracks = 0, evals = 0;
  double E0 = 0.0, Ek = 0.0, last_rk = 0.0, inner_tol = 0.0;
  int which = 0;
  int k = 0, inner_i = 0, ls_j = 0;
  double merit[5] = {0, 0, 0, 0, 0};
  double Eo = 0.0, Ei_c = 0.0, Eo_c = 0.0, tstep = 1.0, m0 = 0.0, current_merit = 0.0;
  int phase = P_TOP;
  bool need_eval = true, pick_xi = false;
  xe = xk;

  for (;;) {
    EvalOut e;
    e.Ei = 0.0;
    e.Eo = 0.0;
    if (need_eval) {
      const bool self_bar = (phase == P_TOP) || (phase == P_FINAL);
      const V bar = v_select(self_bar, xe, xk);
      e = w.evaluate(xe, bar, sigma, alpha, &ri);
      evals++;
    }
    need_eval = true;
    if (phase == P_TOP) {
      Ek = e.Eo;
      last_rk = Ek;
      bool bad = false;
      if (k == 0) {
        E0 = Ek;
        inner_tol = saturate(E0, o.inner_tol_min, o.inner_tol_max, &bad);
      }
      if (!bad && (Ek <= combo_tol || dx_norm <= o.stall_tol)) {
        eflag = FBSTAB_SUCCESS;
        which = 0;
        break;
      }
      if (!bad) inner_tol = saturate(inner_tol * o.delta, o.inner_tol_min, Ek, &bad);
      if (bad) {
        status = FBSTAB_STATUS_SATURATE;
        break;
      }
      xi = xk;
#pragma unroll
      for (int m = 0; m < 5; m++) merit[m] = 0.0;
      Ei_c = e.Ei;
      Eo_c = e.Eo;
      inner_i = 0;
    } else if (phase == P_TRIAL) {
      const double mp = 0.5 * e.Ei * e.Ei;
      if (mp <= m0 - 2.0 * tstep * o.eta * current_merit) {
        xi = xp;
        Ei_c = e.Ei;
        Eo_c = e.Eo;
        inner_i++;
      } else {
        tstep *= o.beta;
        backtracks++;
        ls_j++;
        if (ls_j < o.max_linesearch_iters) {
          v_axpy(w, xi, tstep, dx, &xp);
          xe = xp;
          continue;
        }
        v_axpy(w, xi, tstep, dx, &xi);
        inner_i++;
        if (inner_i < o.max_inner_iters) {
          xe = xi;
          phase = P_REEVAL;
          continue;
        }
      }
    } else if (phase == P_REEVAL) {
      Ei_c = e.Ei;
      Eo_c = e.Eo;
    } else {  // P_FINAL
      last_rk = e.Eo;
      eflag = FBSTAB_MAXITERATIONS;
      which = pick_xi ? 1 : 0;
      break;
    }

    bool inner_done = (inner_i >= o.max_inner_iters);
    if (!inner_done) {
      const double Ei = Ei_c;
      Eo = Eo_c;
      last_rk = Eo;
      if ((Ei <= inner_tol && Eo < Ek) || (Ei <= o.inner_tol_min)) inner_done = true;
      if (newton >= o.max_newton_iters) inner_done = true;
      if (!inner_done) {
        if (!w.newton_step(xi, xk, sigma, alpha, ri, &dx)) {
          status = FBSTAB_STATUS_FACTOR_FAILED;
          which = 0;
          break;
        }
        newton++;
        current_merit = 0.5 * Ei * Ei;
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
        v_axpy(w, xi, tstep, dx, &xp);
        xe = xp;
        phase = P_TRIAL;
        continue;
      }
    }
    xi.v = fmax(xi.v, 0.0);  // ProjectDuals
    if (newton >= o.max_newton_iters) {
      pick_xi = Eo < Ek;
      xe = v_select(pick_xi, xi, xk);
      phase = P_FINAL;
      continue;
    }
    {  // dx = xi - xk (y-aware)
      dx.z = xi.z + (-1.0) * xk.z;
      dx.l = xi.l + (-1.0) * xk.l;
      dx.v = xi.v + (-1.0) * xk.v;
      const double y = xi.y + (-1.0) * xk.y;
      dx.y = y + w.br;
      double sq = (w.half == 0) ? fma(dx.l, dx.l, dx.z * dx.z) : 0.0;
      sq = fma(dx.v, dx.v, sq);
      double2 t0, t1;
      w.sync();
      w.combine2(0, warp_sum(sq), 0.0, &t0, &t1);
      dx_norm = sqrt(t0.x + t1.x);
    }
    if (o.check_feasibility) {
      const int feas = w.feasibility(dx, o.infeas_tol);
      if (feas != 0) {
        eflag = (feas == 1)   ? FBSTAB_PRIMAL_INFEASIBLE
                : (feas == 2) ? FBSTAB_DUAL_INFEASIBLE
                              : FBSTAB_PRIMAL_DUAL_INFEASIBLE;
        which = 2;
        break;
      }
    }
    xk = xi;
    prox++;
    k++;
    if (k >= o.max_prox_iters) {
      eflag = FBSTAB_MAXITERATIONS;
      which = 0;
      break;
    }
    xe = xk;
    phase = P_TOP;
  }

  V r;
  r.z = (which == 0) ? xk.z : (which == 1) ? xi.z : dx.z;
  r.l = (which == 0) ? xk.l : (which == 1) ? xi.l : dx.l;
  r.v = (which == 0) ? xk.v : (which == 1) ? xi.v : dx.v;
  r.y = (which == 0) ? xk.y : (which == 1) ? xi.y : dx.y;
  if (w.half == 0) {
    if (lane < w.nz) A.z[(size_t)inst * w.nz + lane] = r.z;
    if (lane < w.nl) A.l[(size_t)inst * w.nl + lane] = r.l;
  }
  if (tid < w.nv) {
    A.v[(size_t)inst * w.nv + tid] = r.v;
    A.y[(size_t)inst * w.nv + tid] = r.y;
  }
  if (tid == 0) {
    fbstab_out* out = A.out + inst;
    out->eflag = eflag;
    out->newton_iters = newton;
    out->prox_iters = prox;
    out->status = status;
    out->residual = last_rk;
    out->initial_residual = E0;
    out->solve_time = -1.0;
    out->ls_backtracks = backtracks;
    out->residual_evals = evals;
  }
}

// One engine stage on caller-supplied iterates (per-kernel parity tests).
__device__ __forceinline__ void run_component(Team& w, const Args& A, int inst) {
  const fbstab_component_io& io = A.io;
  const int lane = w.lane, tid = w.tid;
  const size_t oz = (size_t)inst * w.nz, ol = (size_t)inst * w.nl, ov = (size_t)inst * w.nv;
  const double alpha = A.opts.alpha;
  V x, xb, dx;
  R ri;
  const bool in = tid < w.nv;
  x.z = (lane < w.nz) ? io.z[oz + lane] : 0.0;
  x.l = (lane < w.nl && io.l) ? io.l[ol + lane] : 0.0;
  xb.z = (lane < w.nz) ? (io.zbar ? io.zbar[oz + lane] : x.z) : 0.0;
  xb.l = (lane < w.nl) ? (io.lbar ? io.lbar[ol + lane] : x.l) : 0.0;
  x.v = (in && io.v) ? io.v[ov + tid] : 0.0;
  x.y = (in && io.y) ? io.y[ov + tid] : 0.0;
  xb.v = in ? (io.vbar ? io.vbar[ov + tid] : x.v) : 0.0;
  xb.y = 0.0;
  if (A.comp == FBSTAB_COMP_MARGIN) {
    w.publish(x);
    const double az = w.Az1();
    if (in) io.dy[ov + tid] = w.br - az;
  } else if (A.comp == FBSTAB_COMP_RESIDUAL) {
    EvalOut e = w.evaluate(x, xb, io.sigma, alpha, &ri);
    const double nzr = ri.z - io.sigma * (x.z - xb.z);
    const double nlr = ri.l - io.sigma * (x.l - xb.l);
    const double n = pnr(x.y, x.v, alpha);
    // z- and l-type parts are replicated: warp 0 reduces them; v-type per warp, combined
    double sq[6];
    sq[0] = warp_sum(ri.z * ri.z);
    sq[1] = warp_sum(ri.l * ri.l);
    sq[3] = warp_sum(nzr * nzr);
    sq[4] = warp_sum(nlr * nlr);
    double2 t0, t1;
    w.sync();
    w.combine2(0, warp_sum(ri.v * ri.v), warp_sum(n * n), &t0, &t1);
    sq[2] = t0.x + t1.x;
    sq[5] = t0.y + t1.y;
    if (w.half == 0) {
      if (lane < w.nz) io.rz[oz + lane] = ri.z;
      if (lane < w.nl) io.rl[ol + lane] = ri.l;
    }
    if (in) io.rv[ov + tid] = ri.v;
    if (tid == 0 && io.norms) {
#pragma unroll
      for (int q = 0; q < 6; q++) io.norms[(size_t)inst * 8 + q] = sqrt(sq[q]);
      io.norms[(size_t)inst * 8 + 6] = e.Ei;
      io.norms[(size_t)inst * 8 + 7] = e.Eo;
    }
  } else if (A.comp == FBSTAB_COMP_NEWTON) {
    ri.z = (lane < w.nz) ? -io.rz[oz + lane] : 0.0;
    ri.l = (lane < w.nl) ? -io.rl[ol + lane] : 0.0;
    ri.v = in ? -io.rv[ov + tid] : 0.0;
    const bool okk = w.newton_step(x, xb, io.sigma, alpha, ri, &dx);
    if (w.half == 0) {
      if (lane < w.nz) io.dz[oz + lane] = dx.z;
      if (lane < w.nl) io.dl[ol + lane] = dx.l;
    }
    if (in) {
      io.dv[ov + tid] = dx.v;
      io.dy[ov + tid] = dx.y;
      if (io.gamma) io.gamma[ov + tid] = w.gamma;
      if (io.mus) io.mus[ov + tid] = w.mus;
    }
    if (tid == 0 && io.status)
      io.status[inst] = okk ? FBSTAB_STATUS_OK : FBSTAB_STATUS_FACTOR_FAILED;
  } else if (A.comp == FBSTAB_COMP_FEAS) {
    const int feas = w.feasibility(x, io.tol);
    if (tid == 0 && io.status) io.status[inst] = feas;
  }
}

template <bool COMPONENT>
__global__ void __launch_bounds__(64 * kTeams, 1)
dense_small2_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) double smem[];
  Team w;
  const int team = threadIdx.x >> 6;
  w.tid = threadIdx.x & 63;
  w.lane = threadIdx.x & 31;
  w.half = w.tid >> 5;
  w.bar = 1 + team;
  w.sb = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)team * (unsigned)(SLAB * sizeof(double));
  w.nz = a.nz;
  w.nl = a.nl;
  w.nv = a.nv;
  int inst = w.next_instance(a.counter);
  while (inst < a.batch) {
    w.load(a, inst);
    const int next = w.next_instance(a.counter);
    if (next < a.batch) w.prefetch(a, next);
    if (COMPONENT)
      run_component(w, a, inst);
    else
      solve_one(w, a, inst);
    inst = next;
  }
}

}  // namespace small2

int DenseSmall2TeamsPerCta() { return small2::kTeams; }

int DenseSmall2Init(DenseSmallPlan* p) {
  const size_t smem = sizeof(double) * small2::SLAB * small2::kTeams;
  if (cudaFuncSetAttribute(small2::dense_small2_kernel<false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(small2::dense_small2_kernel<true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small2::dense_small2_kernel<false>,
                                                    64 * small2::kTeams, smem) != cudaSuccess ||
      occ < 1) {
    cudaGetLastError();
    return 1;
  }
  p->smem2 = smem;
  return 0;
}

int DenseSmall2Launch(const DenseSmallPlan& p, int batch, const double* H, const double* f,
                      const double* G, const double* h, const double* A, const double* b,
                      double* z, double* l, double* v, double* y, fbstab_out* out,
                      const fbstab_options& opts, int comp, const fbstab_component_io* io,
                      cudaStream_t stream) {
  small2::Args a;
  a.comp = comp;
  if (io)
    a.io = *io;
  else
    memset(&a.io, 0, sizeof(a.io));
  a.nz = p.nz;
  a.nl = p.nl;
  a.nv = p.nv;
  a.batch = batch;
  a.H = H;
  a.f = f;
  a.G = G;
  a.h = h;
  a.A = A;
  a.b = b;
  a.z = z;
  a.l = l;
  a.v = v;
  a.y = y;
  a.out = out;
  a.counter = p.counter;
  a.opts = opts;
  const int ctas = (batch + small2::kTeams - 1) / small2::kTeams;
  const int grid = ctas < p.grid ? ctas : p.grid;
  if (comp < 0)
    small2::dense_small2_kernel<false><<<grid, 64 * small2::kTeams, p.smem2, stream>>>(a);
  else
    small2::dense_small2_kernel<true><<<grid, 64 * small2::kTeams, p.smem2, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

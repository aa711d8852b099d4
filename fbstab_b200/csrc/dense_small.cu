// dense_small.cu -- warp-per-problem FBstab for small dense QPs
// (nz <= 32, nl <= 8, nv <= 64; BASELINE config 2 is exactly 32/8/64).
//
// One warp owns one QP instance for its whole solve:
//  * the problem matrices are staged ONCE into the warp's slab of shared memory
//    (zero padded to 32/8/64): A and G with rows contiguous and a padded
//    leading dimension of 34, so that row reads, column reads and the DMMA
//    fragment reads are all bank-conflict free or at worst 2-way; H as its
//    packed lower triangle (H is symmetric; 4.1 KB instead of 8.5 KB), which is
//    what lets EIGHT slabs -- two warps per SM sub-partition -- fit the 227 KB
//    of an SM.  Every access uses explicit shared-space instructions (LDS/STS);
//  * every iterate / residual vector and f, h, b live in registers: lane j owns
//    entry j of z-type vectors, lanes 0..7 own l, lane k owns rows k and k+32
//    of v-type vectors;
//  * E = H + sigma I + A' Gamma A is accumulated on the FP64 tensor cores
//    (mma.sync.m8n8k4.f64, 160 DMMAs for the 10 lower 8x8 blocks);
//  * the KKT matrix K = [E G'; G -sigma I] is eliminated in registers: lane i
//    holds the full symmetric row i of E, the G block is held column-wise
//    (lane j owns column j), pivot rows are broadcast through a small shared
//    buffer, and the right-hand side rides along as a ninth augmented column,
//    which fuses the forward substitution into the factorisation;
//  * the 8x8 Schur complement -sigma I - G E^-1 G' is formed with 16 DMMAs and
//    eliminated by lanes 0..7.
// The kernel is persistent: one CTA of kWarps independent warps per SM; a warp
// that finishes an instance pulls the next index from a global atomic counter
// (no host round trips, natural load balance over the 9..28-iteration spread)
// and prefetches that instance's data into L2 while it is still solving the
// current one.
//
// Round 2 measured and rejected (profiles/r2_dense_small_steps.txt,
// profiles/r2_dense_small_team2_ncu.txt): two warps per instance on the same
// slab (commit 783f97a, dense_small2.cu: parity-green, 1.47x SLOWER -- the
// dependent chain of one instance got longer, and there are two instances per
// scheduler either way), an (even, odd) pair of elimination steps per loop trip
// (-2.5%), the pivot reciprocal computed one step ahead (-4.7%, the switch
// FBSTAB_DS_LOOKAHEAD_RCP is kept for re-measurement), software-pipelined
// mat-vec loops (+0.4%, FBSTAB_DS_PIPE_MATVEC).
//
// Follows the same reference code as engine.cuh / dense_problem.cuh
// (fbstab_algorithm-impl.h:113-304, dense_cholesky_solver.cc:32-148,
// full_residual.cc:49-118, full_feasibility.cc:25-88).
#include <cstring>

#include "common.cuh"
#include "dense_small.h"
#include "engine.cuh"

#ifndef FBSTAB_DS_PAIR_STEPS
#define FBSTAB_DS_PAIR_STEPS 1
#endif

namespace fbs {

namespace small {

constexpr int NZ = 32;  // padded sizes
constexpr int NL = 8;
constexpr int NV = 64;
constexpr int NVR = NV / 32;
constexpr int LD = 34;  // leading dimension of Hs/As/Gs rows
constexpr int kWarps = 8;  // independent warps (= instances in flight) per CTA

// shared-memory layout of one warp's slab (in doubles)
constexpr int OFF_H = 0;                 // Hp[i(i+1)/2 + j] = H(i,j), j <= i (packed lower)
constexpr int H_SIZE = NZ * (NZ + 1) / 2;
constexpr int OFF_A = OFF_H + H_SIZE;    // As[j + LD*k] = A(k,j)
constexpr int OFF_G = OFF_A + NV * LD;   // Gs[j + LD*r] = G(r,j)
constexpr int OFF_ZB = OFF_G + NL * LD;  // broadcast copies: z(32) l(8) v(64)
constexpr int OFF_LB = OFF_ZB + NZ;
constexpr int OFF_VB = OFF_LB + NL;
constexpr int OFF_SCR = OFF_VB + NV;  // scratch shared by the phases of a Newton step
constexpr int SCR_SIZE = 392;  // Gamma (64) | transposition (8*LD) | pivot rows | Schur
// pivot-row buffers of the elimination live in the scratch (free at that time):
// 4 x (32 row entries, pivot first + 10 augmented + pad): rows k, k + 1 being applied and
// rows k + 2, k + 3 being published (slot = row & 3)
constexpr int OFF_COL = OFF_SCR;
constexpr int COL_STRIDE = 44;
constexpr int SLAB = OFF_SCR + SCR_SIZE;
static_assert(SLAB % 2 == 0, "slab must keep 16-byte alignment");
static_assert(OFF_A % 2 == 0 && OFF_G % 2 == 0 && OFF_ZB % 2 == 0 &&
                  OFF_VB % 2 == 0 && OFF_COL % 2 == 0 && OFF_SCR % 2 == 0,
              "LDS.128 targets must be 16-byte aligned");
static_assert((size_t)SLAB * kWarps * sizeof(double) <= 232448,
              "slabs must fit the 227 KB of shared memory of one SM");

struct Args {
  int nz, nl, nv, batch;
  const double *H, *f, *G, *h, *A, *b;
  double *z, *l, *v, *y;
  fbstab_out* out;
  int* counter;
  fbstab_options opts;
  int comp;  // < 0: full solve; otherwise one FBSTAB_COMP_* stage
  fbstab_component_io io;
};

struct V {  // primal-dual iterate in registers
  double z, l, v[NVR], y[NVR];
};
struct R {  // residual in registers
  double z, l, v[NVR];
};

// ---- explicit shared-space accessors (32-bit shared byte addresses) ---------
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
// 16-byte store under a predicate: no branch, so the store can sit between the
// DFMAs of an update (only the owning lane writes)
__device__ __forceinline__ void sts2_if(bool p, unsigned a, double x, double y) {
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.s32 q, %3, 0;\n @q st.shared.v2.f64 [%0], {%1,%2};\n}"
      ::"r"(a), "d"(x), "d"(y), "r"((int)p)
      : "memory");
}
__device__ __forceinline__ void sts_if(bool p, unsigned a, double x) {
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}" ::"r"(a),
               "d"(x), "r"((int)p)
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
__device__ __forceinline__ double bcast(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}
// byte address of double index i in a slab at shared address sb
__device__ __forceinline__ constexpr unsigned D(int i) { return 8u * (unsigned)i; }

struct Warp {
  unsigned sb;  // shared-space byte address of this warp's slab
  int lane;
  int nz, nl, nv;
  double fr, hr, br[NVR];  // f(lane), h(lane), b(lane + 32 m)
  // Newton-step state
  double gamma[NVR], mus[NVR], rmu[NVR];  // rmu = RN(1 / mus)
  double a[NZ];      // row `lane` of E, rotated left once per elimination step
  double g[NL + 1];  // column `lane` of G (rows 0..7) and the rhs (row 8)
  double dinv;       // 1 / pivot of row `lane`
  bool ok;

  // ---- products -------------------------------------------------------------
  __device__ __forceinline__ void publish(const V& x) {
    __syncwarp();
    sts(sb + D(OFF_ZB + lane), x.z);
    if (lane < NL) sts(sb + D(OFF_LB + lane), x.l);
#pragma unroll
    for (int m = 0; m < NVR; m++) sts(sb + D(OFF_VB + lane + 32 * m), x.v[m]);
    __syncwarp();
  }
  // H(i,j) from the packed lower triangle (H(j,i) above the diagonal)
  __device__ __forceinline__ double Hel(int i, int j) const {
    const int r = max(i, j), c = min(i, j);
    return lds(sb + D(OFF_H + (r * (r + 1) >> 1) + c));
  }
  // (A' vb)[lane] and, with WITH_H, (H zb)[lane] in ONE rolled loop of 16 trips
  // (4 rows of A and 2 columns of H per trip): six independent accumulation
  // chains hide the DFMA latency, and the loop body -- not an unrolled copy per
  // call site -- is what the instruction cache holds.  Each dot product keeps
  // the partial-sum order of round 1 ((s0 + s1) + (s2 + s3), even / odd), so
  // the results are bit-identical.
  template <bool WITH_H>
  __device__ __forceinline__ double ATv_Hz(double* hz) const {
    const unsigned ac = sb + D(OFF_A + lane);
    const unsigned vb = sb + D(OFF_VB);
    const unsigned zb = sb + D(OFF_ZB);
    const unsigned hrow = sb + D(OFF_H + (lane * (lane + 1) >> 1));  // H(lane, 0..lane)
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, h0 = 0.0, h1 = 0.0;
#ifdef FBSTAB_DS_PIPE_MATVEC
    // software pipelined: the operands of trip t + 1 are requested before the FMAs of
    // trip t issue (the last trip re-reads trip 0: no branch in the body)
    double2 v0, v1, zz = {0.0, 0.0};
    double a0, a1, a2, a3, e0 = 0.0, e1 = 0.0;
    auto fetch = [&](int t) {
      const int k = 4 * t, j = 2 * t;
      v0 = lds2(vb + D(k));
      v1 = lds2(vb + D(k + 2));
      a0 = lds(ac + D(LD * k));
      a1 = lds(ac + D(LD * (k + 1)));
      a2 = lds(ac + D(LD * (k + 2)));
      a3 = lds(ac + D(LD * (k + 3)));
      if (WITH_H) {
        zz = lds2(zb + D(j));
        e0 = lds(j <= lane ? hrow + D(j) : sb + D(OFF_H + (j * (j + 1) >> 1) + lane));
        e1 = lds(j + 1 <= lane ? hrow + D(j + 1)
                               : sb + D(OFF_H + ((j + 1) * (j + 2) >> 1) + lane));
      }
    };
    fetch(0);
#pragma unroll 2
    for (int t = 0; t < NV / 4; t++) {
      const double2 cv0 = v0, cv1 = v1, czz = zz;
      const double c0 = a0, c1 = a1, c2 = a2, c3 = a3, ce0 = e0, ce1 = e1;
      fetch((t + 1) & (NV / 4 - 1));
      if (WITH_H) {
        h0 = fma(ce0, czz.x, h0);
        h1 = fma(ce1, czz.y, h1);
      }
      s0 = fma(c0, cv0.x, s0);
      s1 = fma(c1, cv0.y, s1);
      s2 = fma(c2, cv1.x, s2);
      s3 = fma(c3, cv1.y, s3);
    }
#else
#pragma unroll 2
    for (int t = 0; t < NV / 4; t++) {
      const int k = 4 * t, j = 2 * t;
      const double2 v0 = lds2(vb + D(k));
      const double2 v1 = lds2(vb + D(k + 2));
      const double a0 = lds(ac + D(LD * k)), a1 = lds(ac + D(LD * (k + 1)));
      const double a2 = lds(ac + D(LD * (k + 2))), a3 = lds(ac + D(LD * (k + 3)));
      if (WITH_H) {
        const double2 zz = lds2(zb + D(j));
        // column part: H(j, lane) for j > lane sits at j(j+1)/2 + lane
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
#endif
    if (WITH_H) *hz = h0 + h1;
    return (s0 + s1) + (s2 + s3);
  }
  // (A zb)[lane + 32 m]: 2 NVR chains, rolled
  __device__ __forceinline__ void Az(double (&o)[NVR]) const {
    const unsigned zb = sb + D(OFF_ZB);
    const unsigned arow = sb + D(OFF_A + LD * lane);
    double t0[NVR], t1[NVR];
#pragma unroll
    for (int m = 0; m < NVR; m++) t0[m] = t1[m] = 0.0;
#pragma unroll 4
    for (int j = 0; j < NZ; j += 2) {
      const double2 zz = lds2(zb + D(j));
#pragma unroll
      for (int m = 0; m < NVR; m++) {
        const double2 aa = lds2(arow + D(LD * 32 * m + j));
        t0[m] = fma(aa.x, zz.x, t0[m]);
        t1[m] = fma(aa.y, zz.y, t1[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < NVR; m++) o[m] = t0[m] + t1[m];
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
  // (G zb)[lane % 8], same value in the four lanes sharing lane % 8
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
  // ---- fused residual evaluation (see engine.cuh) ---------------------------
  __device__ __forceinline__ EvalOut evaluate(const V& x, const V& xbar,
                                              double sigma, double alpha, R* ri) {
    publish(x);
    double hz;
    const double atv = ATv_Hz<true>(&hz);
    double tz = fr + hz;
    tz += GTl();
    tz += atv;
    const double gz = Gz();
    const double tl = (lane < NL) ? hr - gz : 0.0;
    double si = 0.0, so = 0.0;
    {
      const double r = tz + sigma * (x.z - xbar.z);
      ri->z = r;
      si = r * r;
      so = tz * tz;
    }
    {
      const double r = tl + sigma * (x.l - xbar.l);
      ri->l = r;
      si = fma(r, r, si);
      so = fma(tl, tl, so);
    }
#pragma unroll
    for (int m = 0; m < NVR; m++) {
      const double y = x.y[m], v = x.v[m];
      const double ys = y + sigma * (v - xbar.v[m]);
      const double r = pfb(ys, v, alpha);
      ri->v[m] = r;
      si = fma(r, r, si);
      const double n = pnr(y, v, alpha);
      so = fma(n, n, so);
    }
    EvalOut e;
    e.Ei = sqrt(warp_sum(si));
    e.Eo = sqrt(warp_sum(so));
    return e;
  }

  // Gauss-Jordan elimination, one segment: steps k0..k1-1, during which at most
  // W trailing columns are still non-zero.  Lane `row` owns row `row` of the
  // matrix in a[0..W] and NA augmented entries in g[]; every step the row
  // registers rotate left by one, so a[0] is always the entry in the pivot
  // column and the loop body is the same for every k.  The loop is kept ROLLED
  // on purpose: the CTA's warps run different instances at different program
  // counters, and a fully unrolled elimination measured 43% of all stall
  // samples on instruction fetch (profiles/r1_dense_small_unrolled_elimination.txt).
  //
  // The pivot row travels through shared memory (two buffers, one __syncwarp
  // per step).  It is the pivot row ITSELF, not the pivot column taken from
  // the other lanes: after the 1e8-scale cancellations of A' Gamma A the (k,j)
  // and (j,k) entries differ at the 1e-8 relative level, and mixing them ruins
  // the accuracy of dz along the active-constraint normals.
  //
  // Round 2 (profiles/r2_dense_small_steps.txt): the owner of the NEXT pivot row
  // stores it pair by pair as the update produces it -- predicated stores between
  // the DFMAs -- instead of in a 21-store block at the top of the next step
  // (210 + 66 cycles of store issue and drain on the critical path of every
  // step); pairs (a[0],a[1]), (a[2],a[3]), ... are register-aligned, so each
  // 16-byte store takes its operands in place (the (a[1],a[2]) pairing of round
  // 1 cost three register moves per store); the row is loaded in chunks of four
  // pairs, one chunk ahead of the update (the whole row in flight cost 84
  // registers and pushed loop invariants out of the register file).  Same
  // operations in the same order: results are bit-identical to round 1.
  // Precondition: row k0 is already in buffer k0 & 1 (store_row / previous segment).
  template <int W, int NA, int NG>
  __device__ __forceinline__ void eliminate_step(int k, unsigned cb, unsigned nb, int row,
                                                 double (&a)[NZ], double (&g)[NG],
                                                 double* rpiv) {
    static_assert(W & 1, "a[0..W] travels as (W + 1) / 2 aligned pairs");
    constexpr int NP = (W + 1) / 2, NAP = (NA + 1) / 2, NCH = (NP + 3) / 4;
    __syncwarp();
    double2 c[2][4], xr[NAP];
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < NP) c[0][p] = lds2(cb + D(2 * p));
#pragma unroll
    for (int r = 0; r < NAP; r++) xr[r] = lds2(cb + D(32 + 2 * r));
    const double d = c[0][0].x;
    if (!(fabs(d) > 0.0)) ok = false;
#ifdef FBSTAB_DS_LOOKAHEAD_RCP
    // the pivot's reciprocal was computed one step ahead by the row's owner, off the
    // store -> barrier -> load -> reciprocal -> multiply chain of a step
    const double rd = lds(cb + D(42));
#else
    const double rd = 1.0 / d;
#endif
    if (row == k) *rpiv = rd;
    const double lik = (row != k) ? a[0] * rd : 0.0;
    const bool nxt = (row == k + 1);
#pragma unroll
    for (int q = 0; q < NCH; q++) {
#pragma unroll
      for (int p = 0; p < 4; p++)
        if (4 * (q + 1) + p < NP) c[(q + 1) & 1][p] = lds2(cb + D(8 * (q + 1) + 2 * p));
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int pg = 4 * q + p;  // pair (entries 2 pg, 2 pg + 1) of the pivot row
        if (pg < NP) {
          if (pg >= 1) {
            a[2 * pg - 1] = fma(-lik, c[q & 1][p].x, a[2 * pg]);
            sts2_if(nxt, nb + D(2 * pg - 2), a[2 * pg - 2], a[2 * pg - 1]);
          }
          a[2 * pg] = fma(-lik, c[q & 1][p].y, a[2 * pg + 1]);
        }
      }
    }
    a[W] = 0.0;
    sts2_if(nxt, nb + D(W - 1), a[W - 1], 0.0);
#ifdef FBSTAB_DS_LOOKAHEAD_RCP
    // a[0] is final since the first FMA of the update: the division overlaps the rest
    // (other lanes divide 1 by 1: the fast path, whatever their entry is)
    sts_if(nxt, nb + D(42), 1.0 / (nxt ? a[0] : 1.0));
#endif
#pragma unroll
    for (int r = 0; r < NA; r++) {
      const double xm = (r & 1) ? xr[r / 2].y : xr[r / 2].x;
      g[r] = fma(-lik, xm, g[r]);
      if (r & 1) sts2_if(nxt, nb + D(32 + r - 1), g[r - 1], g[r]);
    }
    if (NA & 1) sts2_if(nxt, nb + D(32 + NA - 1), g[NA - 1], 0.0);
  }
  // Round 2, second pass: TWO pivots per trip.  A step is a chain -- pivot-row store ->
  // __syncwarp -> broadcast load -> reciprocal -> multiplier -> update -> next store --
  // and the warp sits in it 32 times per Newton step with one other warp to hide it.
  // Here the rows of pivots k AND k + 1 (the second one as it is BEFORE pivot k's update)
  // are published together, every lane redoes row k + 1's update by pivot k on the fly
  // (r1_j = fma(-m, p_j, q_j): the very FMAs its owner executes, so the values are the
  // owner's bit for bit) and applies both pivots to its own row: half the barriers and
  // store -> load round trips for one redundant FMA per entry.  Same operations in the
  // same order on every entry: results are bit-identical to the one-pivot steps
  // (FBSTAB_DS_PAIR_STEPS=0 keeps those for A/B).
  // Precondition: rows k0 and k0 + 1 are in slots (k0 & 3), (k0 + 1) & 3, both aligned on
  // column k0 (store_rows2 / previous trip).
  template <int W, int NA, int NG>
  __device__ __forceinline__ void eliminate_step2(int k, unsigned pb, unsigned qb, unsigned n0,
                                                  unsigned n1, int row, double (&a)[NZ],
                                                  double (&g)[NG], double* rpiv) {
    static_assert(W & 1, "a[0..W] travels as (W + 1) / 2 aligned pairs");
    constexpr int NP = (W + 1) / 2, NAP = (NA + 1) / 2, CH = 2, NCH = (NP + CH - 1) / CH;
    __syncwarp();
    double2 cp[2][CH], cq[2][CH], xp[NAP], xq[NAP];
#pragma unroll
    for (int p = 0; p < CH; p++)
      if (p < NP) {
        cp[0][p] = lds2(pb + D(2 * p));
        cq[0][p] = lds2(qb + D(2 * p));
      }
#pragma unroll
    for (int r = 0; r < NAP; r++) {
      xp[r] = lds2(pb + D(32 + 2 * r));
      xq[r] = lds2(qb + D(32 + 2 * r));
    }
    const double d = cp[0][0].x;
    if (!(fabs(d) > 0.0)) ok = false;
    const double rd = 1.0 / d;
    const double m = cq[0][0].x * rd;                    // row k + 1's multiplier for pivot k
    const double d2 = fma(-m, cp[0][0].y, cq[0][0].y);   // pivot k + 1
    if (!(fabs(d2) > 0.0)) ok = false;
    const double rd2 = 1.0 / d2;
    if (row == k) *rpiv = rd;
    if (row == k + 1) *rpiv = rd2;
    const double lik = (row != k) ? a[0] * rd : 0.0;
    const double t1 = fma(-lik, cp[0][0].y, a[1]);
    const double l2 = (row != k + 1) ? t1 * rd2 : 0.0;
    const bool st = (row == k + 2) || (row == k + 3);
    const unsigned nb = (row == k + 3) ? n1 : n0;
#pragma unroll
    for (int q = 0; q < NCH; q++) {
#pragma unroll
      for (int p = 0; p < CH; p++)
        if (CH * (q + 1) + p < NP) {
          cp[(q + 1) & 1][p] = lds2(pb + D(2 * (CH * (q + 1) + p)));
          cq[(q + 1) & 1][p] = lds2(qb + D(2 * (CH * (q + 1) + p)));
        }
#pragma unroll
      for (int p = 0; p < CH; p++) {
        const int pg = CH * q + p;  // pair (entries 2 pg, 2 pg + 1) of the pivot rows
        if (pg >= 1 && pg < NP) {
          const double2 pp = cp[q & 1][p], qq = cq[q & 1][p];
          const double tx = fma(-lik, pp.x, a[2 * pg]), ty = fma(-lik, pp.y, a[2 * pg + 1]);
          const double rx = fma(-m, pp.x, qq.x), ry = fma(-m, pp.y, qq.y);
          a[2 * pg - 2] = fma(-l2, rx, tx);
          a[2 * pg - 1] = fma(-l2, ry, ty);
          sts2_if(st, nb + D(2 * pg - 2), a[2 * pg - 2], a[2 * pg - 1]);
        }
      }
    }
    a[W - 1] = 0.0;
    a[W] = 0.0;
    sts2_if(st, nb + D(W - 1), 0.0, 0.0);
#pragma unroll
    for (int r = 0; r < NA; r++) {
      const double mp = (r & 1) ? xp[r / 2].y : xp[r / 2].x;
      const double mq = (r & 1) ? xq[r / 2].y : xq[r / 2].x;
      g[r] = fma(-l2, fma(-m, mp, mq), fma(-lik, mp, g[r]));
      if (r & 1) sts2_if(st, nb + D(32 + r - 1), g[r - 1], g[r]);
    }
    if (NA & 1) sts2_if(st, nb + D(32 + NA - 1), g[NA - 1], 0.0);
  }
  // (An (even, odd) pair of steps per trip -- loop-invariant buffer addresses instead of
  // rewriting the address register of stores still in flight -- measured 2.5% SLOWER:
  // the body no longer fits the instruction cache next to its neighbours.)
  template <int W, int NA, int NG>
  __device__ __forceinline__ void eliminate(int k0, int k1, int row, double (&a)[NZ],
                                            double (&g)[NG], double* rpiv) {
#if FBSTAB_DS_PAIR_STEPS
#pragma unroll 1
    for (int k = k0; k < k1; k += 2) {
      const unsigned base = sb + D(OFF_COL);
      eliminate_step2<W, NA>(k, base + D(COL_STRIDE * (k & 3)),
                             base + D(COL_STRIDE * ((k + 1) & 3)),
                             base + D(COL_STRIDE * ((k + 2) & 3)),
                             base + D(COL_STRIDE * ((k + 3) & 3)), row, a, g, rpiv);
    }
#else
#pragma unroll 1
    for (int k = k0; k < k1; k++) {
      const unsigned cb = sb + D(OFF_COL + COL_STRIDE * (k & 1));
      const unsigned nb = sb + D(OFF_COL + COL_STRIDE * ((k + 1) & 1));
      eliminate_step<W, NA>(k, cb, nb, row, a, g, rpiv);
    }
#endif
  }
  // Row `row` == k0 into pivot buffer k0 & 1 (before the first segment).
  template <int W, int NA, int NG>
  __device__ __forceinline__ void store_row(int k0, int row, const double (&a)[NZ],
                                            const double (&g)[NG]) {
#if FBSTAB_DS_PAIR_STEPS
    // rows k0 and k0 + 1, each into its slot (row & 3), both aligned on column k0
    const unsigned cb = sb + D(OFF_COL + COL_STRIDE * (row & 3));
    const bool own = (row == k0) || (row == k0 + 1);
#else
    const unsigned cb = sb + D(OFF_COL + COL_STRIDE * (k0 & 1));
    const bool own = (row == k0);
#endif
#pragma unroll
    for (int m = 0; m <= W; m += 2) sts2_if(own, cb + D(m), a[m], a[m + 1]);
#pragma unroll
    for (int r = 0; r < NA; r += 2)
      sts2_if(own, cb + D(32 + r), g[r], (r + 1 < NA) ? g[r + 1] : 0.0);
#ifdef FBSTAB_DS_LOOKAHEAD_RCP
    sts_if(own, cb + D(42), 1.0 / (own ? a[0] : 1.0));
#endif
  }

  // ---- Newton step: LinearSolver::Initialize + ::Solve fused -----------------
  // Solves V(x,xbar,sigma) dx = -ri.  Returns false on a zero / NaN pivot.
  __device__ __forceinline__ bool newton_step(const V& x, const V& xbar,
                                              double sigma, double alpha,
                                              const R& ri, V* dx) {
    const int r8 = lane >> 2, c4 = lane & 3;
    const unsigned scr = sb + D(OFF_SCR);
    ok = true;
    {
      double r2[NVR], Gam[NVR];
#pragma unroll
      for (int m = 0; m < NVR; m++) {
        const double ys = x.y[m] + sigma * (x.v[m] - xbar.v[m]);
        pfb_barrier(ys, x.v[m], alpha, sigma, &gamma[m], &mus[m]);
        rmu[m] = 1.0 / mus[m];
        Gam[m] = div_r(gamma[m], mus[m], rmu[m]);
        r2[m] = div_r(-ri.v[m], mus[m], rmu[m]);
      }
      // r1z = -rz - A'(rv/mus) ; Gamma -> shared for the DMMA operand scaling
      __syncwarp();
#pragma unroll
      for (int m = 0; m < NVR; m++) {
        sts(sb + D(OFF_VB + lane + 32 * m), r2[m]);
        sts(scr + D(lane + 32 * m), Gam[m]);
      }
      __syncwarp();
    }
    g[NL] = (-ri.z) - ATv_Hz<false>(nullptr);

    // E (lower 8x8 blocks) on the FP64 tensor cores
    {
      double C[4][4][2];
#pragma unroll
      for (int I = 0; I < 4; I++)
#pragma unroll
        for (int J = 0; J <= I; J++) {
          const int hr = 8 * I + r8, hc = 8 * J + 2 * c4;
          C[I][J][0] = Hel(hr, hc) + ((I == J && r8 == 2 * c4) ? sigma : 0.0);
          C[I][J][1] = Hel(hr, hc + 1) + ((I == J && r8 == 2 * c4 + 1) ? sigma : 0.0);
        }
      // operands of chunk kc + 1 are requested before the DMMAs of chunk kc issue
      const unsigned ar0 = sb + D(OFF_A + LD * c4 + r8);
      const unsigned gk0 = scr + D(c4);
      double an[4], gn;
      gn = lds(gk0);
#pragma unroll
      for (int X = 0; X < 4; X++) an[X] = lds(ar0 + D(8 * X));
#pragma unroll 1
      for (int kc = 0; kc < NV / 4; kc++) {
        double af[4], bf[4];
#pragma unroll
        for (int X = 0; X < 4; X++) {
          af[X] = an[X];
          bf[X] = gn * an[X];
        }
        // (the last trip re-reads chunk 0: no branch in the loop body)
        const int kn = (kc + 1) & (NV / 4 - 1);
        gn = lds(gk0 + D(4 * kn));
#pragma unroll
        for (int X = 0; X < 4; X++) an[X] = lds(ar0 + D(LD * 4 * kn + 8 * X));
#pragma unroll
        for (int I = 0; I < 4; I++)
#pragma unroll
          for (int J = 0; J <= I; J++) dmma(C[I][J][0], C[I][J][1], af[I], bf[J]);
      }
      // fragments -> full symmetric rows: lane i gets a[j] = E(i,j) for all j
#pragma unroll
      for (int I = 0; I < 4; I++) {
        __syncwarp();
#pragma unroll
        for (int J = 0; J <= I; J++)
          sts2(scr + D(LD * r8 + 8 * J + 2 * c4), C[I][J][0], C[I][J][1]);
        __syncwarp();
        if ((lane >> 3) == I) {
          const unsigned row = scr + D(LD * (lane & 7));
#pragma unroll
          for (int j = 0; j < NZ; j += 2) {
            if (j < 8 * (I + 1)) {
              const double2 t = lds2(row + D(j));
              a[j] = t.x;
              a[j + 1] = t.y;
            }
          }
        } else if ((lane >> 3) < I) {
#pragma unroll
          for (int rr = 0; rr < 8; rr++) a[8 * I + rr] = lds(scr + D(LD * rr + lane));
        }
      }
    }
    // G block column-wise (lane j owns column j); the rhs is already in g[8]
#pragma unroll
    for (int r = 0; r < NL; r++) g[r] = lds(sb + D(OFF_G + LD * r + lane));

    // Gauss-Jordan elimination of the E block; the
    // pivot-row buffers alias the transposition scratch read just above
    __syncwarp();
    store_row<31, NL + 1>(0, lane, a, g);
    eliminate<31, NL + 1>(0, 8, lane, a, g, &dinv);
    eliminate<23, NL + 1>(8, 16, lane, a, g, &dinv);
    eliminate<15, NL + 1>(16, 24, lane, a, g, &dinv);
    eliminate<7, NL + 1>(24, 32, lane, a, g, &dinv);
    // lane i now holds d_i * (E^-1 [G' a])(i,:) in g[0..8] and dinv = 1/d_i

    // Schur complement S = -sigma I - G Y, rhs c - G t  with [Y t] = E^-1 [G' a]
    __syncwarp();
    {
      // Ys[r + 10*i], r = 0..8
#pragma unroll
      for (int r = 0; r < NL + 1; r += 2)
        sts2(scr + D(10 * lane + r), g[r] * dinv,
             (r + 1 < NL + 1) ? g[r + 1] * dinv : 0.0);
    }
    __syncwarp();
    double S0 = 0.0, S1 = 0.0, T0 = 0.0, T1 = 0.0;
#pragma unroll 2
    for (int kc = 0; kc < NZ / 4; kc++) {
      const int i = 4 * kc + c4;
      const double ge = lds(sb + D(OFF_G + LD * r8 + i));  // G(r8, i)
      const double ye = lds(scr + D(10 * i + r8));         // Y(i, r8)
      const double te = lds(scr + D(10 * i + NL));         // t(i)
      dmma(S0, S1, ge, ye);
      dmma(T0, T1, ge, te);
    }
    __syncwarp();
    sts2(scr + D(320 + 8 * r8 + 2 * c4), S0, S1);
    if (c4 == 0) sts(scr + D(384 + r8), T0);
    __syncwarp();
    double dl = 0.0;
    {
      // 8x8 Gauss-Jordan elimination with the same routine: lane r (and its three
      // replicas r + 8, r + 16, r + 24) owns row r in a[0..7] -- the E rows are
      // used up -- and the right-hand side rides along as the one augmented entry
      const int r = lane & 7;
#pragma unroll
      for (int j = 0; j < NL; j += 2) {
        const double2 t = lds2(scr + D(320 + 8 * r + j));
        a[j] = -t.x - ((j == r) ? sigma : 0.0);
        a[j + 1] = -t.y - ((j + 1 == r) ? sigma : 0.0);
      }
      double sg[2], dsi = 0.0;
      sg[0] = ((lane < NL) ? ri.l : 0.0) - lds(scr + D(384 + r));
      sg[1] = 0.0;
      __syncwarp();  // every lane has read S and T: the pivot buffers may be reused
      // (row index = lane: the replicas in lanes 8..31 never own a pivot row, so
      // they never store; what they compute is not used)
      store_row<NL - 1, 1>(0, lane, a, sg);
      eliminate<NL - 1, 1>(0, NL, lane, a, sg, &dsi);
      dl = (lane < NL) ? sg[0] * dsi : 0.0;
    }
    // dz = t - Y dl = (g[8] - sum_r g[r] dl_r) / d
    double acc = g[NL];
#pragma unroll
    for (int r = 0; r < NL; r++) acc = fma(-g[r], bcast(dl, r), acc);
    const double dz = acc * dinv;
    dx->z = dz;
    dx->l = dl;
    // dv = (rv + gamma .* (A dz)) ./ mus ; dy = b - A dz
    __syncwarp();
    sts(sb + D(OFF_ZB + lane), dz);
    __syncwarp();
    double adz[NVR];
    Az(adz);
#pragma unroll
    for (int m = 0; m < NVR; m++) {
      dx->v[m] = div_r(gamma[m] * adz[m] + (-ri.v[m]), mus[m], rmu[m]);
      dx->y[m] = br[m] - adz[m];
    }
    return ok;
  }

  // FullFeasibility::CheckFeasibility on dx
  __device__ __forceinline__ int feasibility(const V& dx, double tol) {
    publish(dx);
    double adz[NVR];
    Az(adz);
    double d1 = -INFINITY;
#pragma unroll
    for (int m = 0; m < NVR; m++)
      if (lane + 32 * m < nv) d1 = fmax(d1, adz[m]);
    const double gz = Gz();  // shuffles inside: every lane must call it
    const double d2 = (lane < NL) ? fabs(gz) : 0.0;
    double hz;
    const double atv = ATv_Hz<true>(&hz);
    const double d3 = fabs(hz);
    const double w = warp_max(fabs(dx.z));
    const double d4 = warp_sum(fr * dx.z);
    const double p1 = warp_max(fabs(atv + GTl()));
    double p2 = (lane < NL) ? hr * dx.l : 0.0;
    double umax = fabs(dx.l);
#pragma unroll
    for (int m = 0; m < NVR; m++) {
      p2 = fma(br[m], dx.v[m], p2);
      umax = fmax(umax, fabs(dx.v[m]));
    }
    p2 = warp_sum(p2);
    umax = warp_max(umax);
    const double D1 = warp_max(d1), D2 = warp_max(d2), D3 = warp_max(d3);
    const bool dual_inf = (D1 <= w * tol) && (D2 <= tol * w) && (D3 <= tol * w) &&
                          (d4 < 0.0) && (w > 1e-14);
    const bool primal_inf = (p1 <= tol * umax) && (p2 < 0.0);
    return (primal_inf ? 1 : 0) + (dual_inf ? 2 : 0);
  }

  // ---- data staging ----------------------------------------------------------
  // Transposes one column-major rows x cols matrix into the row-contiguous
  // padded layout; loads are issued in batches of 8 per lane so that the
  // global-memory latency is paid once per batch, not once per element.
  __device__ __forceinline__ void stage_matrix(const double* src, int rows, int cols,
                                               int off) {
    const int n = rows * cols;
    for (int e0 = 0; e0 < n; e0 += 256) {
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
  // The lower triangle of the column-major nz x nz matrix H into the packed layout.
  __device__ __forceinline__ void stage_lower(const double* src, int rows) {
    const int n = rows * rows;
    for (int e0 = 0; e0 < n; e0 += 256) {
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
  // Exact 32 / 8 / 64 shape (BASELINE config 2): compile-time index arithmetic
  // (no integer division per element) and 16 loads in flight per lane.
  __device__ __forceinline__ void load_exact(const Args& a_, int inst) {
    const double* Asrc = a_.A + (size_t)inst * NV * NZ;
    const double* Hsrc = a_.H + (size_t)inst * NZ * NZ;
    const double* Gsrc = a_.G + (size_t)inst * NL * NZ;
    // A(r, c) at r + 64 c -> As[c + LD r]; element e = lane + 32 u: c = u / 2, r = lane + 32 (u & 1)
#pragma unroll 1
    for (int u0 = 0; u0 < 2 * NZ; u0 += 16) {
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; u++) t[u] = __ldg(Asrc + lane + 32 * (u0 + u));
      const unsigned base = sb + D(OFF_A + LD * lane + (u0 >> 1));
#pragma unroll
      for (int u = 0; u < 16; u++) sts(base + D(LD * 32 * (u & 1) + (u >> 1)), t[u]);
    }
    // H(r, c) at r + 32 c, lower triangle only -> Hp[r (r + 1) / 2 + c]; e = lane + 32 u: c = u, r = lane
#pragma unroll 1
    for (int u0 = 0; u0 < NZ; u0 += 16) {
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; u++)
        t[u] = (lane >= u0 + u) ? __ldg(Hsrc + lane + 32 * (u0 + u)) : 0.0;
      const unsigned base = sb + D(OFF_H + (lane * (lane + 1) >> 1) + u0);
#pragma unroll
      for (int u = 0; u < 16; u++)
        if (lane >= u0 + u) sts(base + D(u), t[u]);
    }
    // G(r, c) at r + 8 c -> Gs[c + LD r]; e = lane + 32 u: c = lane / 8 + 4 u, r = lane % 8
    {
      double t[NZ / 4];
#pragma unroll
      for (int u = 0; u < NZ / 4; u++) t[u] = __ldg(Gsrc + lane + 32 * u);
      const unsigned base = sb + D(OFF_G + LD * (lane & 7) + (lane >> 3));
#pragma unroll
      for (int u = 0; u < NZ / 4; u++) sts(base + D(4 * u), t[u]);
    }
  }
  __device__ __forceinline__ void load(const Args& a_, int inst) {
    __syncwarp();
    if (nz == NZ && nl == NL && nv == NV) {
      load_exact(a_, inst);
    } else {
      // zero padding of the unused part
      for (int e = 2 * lane; e < OFF_ZB; e += 64) sts2(sb + D(e), 0.0, 0.0);
      __syncwarp();
      stage_lower(a_.H + (size_t)inst * nz * nz, nz);
      stage_matrix(a_.A + (size_t)inst * nv * nz, nv, nz, OFF_A);
      stage_matrix(a_.G + (size_t)inst * nl * nz, nl, nz, OFF_G);
    }
    fr = (lane < nz) ? __ldg(a_.f + (size_t)inst * nz + lane) : 0.0;
    hr = (lane < nl) ? __ldg(a_.h + (size_t)inst * nl + lane) : 0.0;
#pragma unroll
    for (int m = 0; m < NVR; m++)
      br[m] = (lane + 32 * m < nv) ? __ldg(a_.b + (size_t)inst * nv + lane + 32 * m) : 0.0;
    __syncwarp();
  }
  // Pulls instance `inst`'s matrices towards L2 (one 128-byte line per request).
  __device__ __forceinline__ void prefetch(const Args& a_, int inst) const {
    const char* H = (const char*)(a_.H + (size_t)inst * nz * nz);
    const char* A = (const char*)(a_.A + (size_t)inst * nv * nz);
    const char* G = (const char*)(a_.G + (size_t)inst * nl * nz);
    for (int o = 128 * lane; o < 8 * nz * nz; o += 128 * 32) prefetch_l2(H + o);
    for (int o = 128 * lane; o < 8 * nv * nz; o += 128 * 32) prefetch_l2(A + o);
    for (int o = 128 * lane; o < 8 * nl * nz; o += 128 * 32) prefetch_l2(G + o);
  }
};

__device__ __forceinline__ void v_axpy(const Warp& w, const V& src, double a,
                                       const V& dx, V* dst) {
  dst->z = src.z + a * dx.z;
  dst->l = src.l + a * dx.l;
#pragma unroll
  for (int m = 0; m < NVR; m++) {
    dst->v[m] = src.v[m] + a * dx.v[m];
    const double y = src.y[m] + a * dx.y[m];
    dst->y[m] = y + (-a) * w.br[m];
  }
}

__device__ __forceinline__ V v_select(bool c, const V& a, const V& b) {
  V r;
  r.z = c ? a.z : b.z;
  r.l = c ? a.l : b.l;
#pragma unroll
  for (int m = 0; m < NVR; m++) {
    r.v[m] = c ? a.v[m] : b.v[m];
    r.y[m] = c ? a.y[m] : b.y[m];
  }
  return r;
}

// FBstabAlgorithm::Solve for one instance, executed by one warp.  Same
// semantics as engine.cuh::solve_instance; written as a phase machine so that
// the (large) residual evaluation and Newton step each have ONE call site:
// the instruction footprint of the hot loop has to stay inside the
// instruction cache.
enum Phase { P_TOP = 0, P_TRIAL = 1, P_REEVAL = 2, P_FINAL = 3 };

__device__ __forceinline__ void solve_one(Warp& w, const Args& A, int inst) {
  const fbstab_options& o = A.opts;
  const int lane = w.lane;
  const double sigma = o.sigma0, alpha = o.alpha;
  V xk, xi, xp, dx, xe;
  R ri;
  xk.z = (lane < w.nz) ? A.z[(size_t)inst * w.nz + lane] : 0.0;
  xk.l = (lane < w.nl) ? A.l[(size_t)inst * w.nl + lane] : 0.0;
#pragma unroll
  for (int m = 0; m < NVR; m++)
    xk.v[m] = (lane + 32 * m < w.nv) ? A.v[(size_t)inst * w.nv + lane + 32 * m] : 0.0;
  // forcing norm and margin y = b - A z
  double fn = w.fr * w.fr;
  if (lane < NL) fn = fma(w.hr, w.hr, fn);
#pragma unroll
  for (int m = 0; m < NVR; m++) fn = fma(w.br[m], w.br[m], fn);
  const double combo_tol = o.abs_tol + o.rel_tol * (1.0 + sqrt(warp_sum(fn)));
  w.publish(xk);
  {
    double az[NVR];
    w.Az(az);
#pragma unroll
    for (int m = 0; m < NVR; m++) xk.y[m] = w.br[m] - az[m];
  }
  xi = xk;
  xp = xk;
  dx = xk;
  double dx_norm = sqrt((double)w.nz + (double)w.nl + (double)w.nv);
  int eflag = FBSTAB_MAXITERATIONS, status = FBSTAB_STATUS_OK;
  int newton = 0, prox = 0, backtracks = 0, evals = 0;
  double E0 = 0.0, Ek = 0.0, last_rk = 0.0, inner_tol = 0.0;
  int which = 0;  // 0: xk, 1: xi, 2: dx is the result
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
      // top of the proximal loop, fbstab_algorithm-impl.h:158-185
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
      // Armijo test, impl:286-296
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
        // every trial failed: the step is still taken (impl:295-298)
        v_axpy(w, xi, tstep, dx, &xi);
        inner_i++;
        if (inner_i < o.max_inner_iters) {
          xe = xi;
          phase = P_REEVAL;
          continue;
        }
        // falls straight into "inner loop exhausted"
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

    // ---- top of an inner (Newton) iteration, impl:237-260
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
    // ---- subproblem finished, impl:300-216
#pragma unroll
    for (int m = 0; m < NVR; m++) xi.v[m] = fmax(xi.v[m], 0.0);  // ProjectDuals
    if (newton >= o.max_newton_iters) {  // impl:188-199
      pick_xi = Eo < Ek;
      xe = v_select(pick_xi, xi, xk);
      phase = P_FINAL;
      continue;
    }
    {  // dx = xi - xk (y-aware)
      dx.z = xi.z + (-1.0) * xk.z;
      dx.l = xi.l + (-1.0) * xk.l;
      double sq = dx.z * dx.z;
      sq = fma(dx.l, dx.l, sq);
#pragma unroll
      for (int m = 0; m < NVR; m++) {
        dx.v[m] = xi.v[m] + (-1.0) * xk.v[m];
        sq = fma(dx.v[m], dx.v[m], sq);
        const double y = xi.y[m] + (-1.0) * xk.y[m];
        dx.y[m] = y + w.br[m];
      }
      dx_norm = sqrt(warp_sum(sq));
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
#pragma unroll
  for (int m = 0; m < NVR; m++) {
    r.v[m] = (which == 0) ? xk.v[m] : (which == 1) ? xi.v[m] : dx.v[m];
    r.y[m] = (which == 0) ? xk.y[m] : (which == 1) ? xi.y[m] : dx.y[m];
  }
  if (lane < w.nz) A.z[(size_t)inst * w.nz + lane] = r.z;
  if (lane < w.nl) A.l[(size_t)inst * w.nl + lane] = r.l;
#pragma unroll
  for (int m = 0; m < NVR; m++)
    if (lane + 32 * m < w.nv) {
      A.v[(size_t)inst * w.nv + lane + 32 * m] = r.v[m];
      A.y[(size_t)inst * w.nv + lane + 32 * m] = r.y[m];
    }
  if (lane == 0) {
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
__device__ __forceinline__ void run_component(Warp& w, const Args& A, int inst) {
  const fbstab_component_io& io = A.io;
  const int lane = w.lane;
  const size_t oz = (size_t)inst * w.nz, ol = (size_t)inst * w.nl,
               ov = (size_t)inst * w.nv;
  const double alpha = A.opts.alpha;
  V x, xb, dx;
  R ri;
  x.z = (lane < w.nz) ? io.z[oz + lane] : 0.0;
  x.l = (lane < w.nl && io.l) ? io.l[ol + lane] : 0.0;
  xb.z = (lane < w.nz) ? (io.zbar ? io.zbar[oz + lane] : x.z) : 0.0;
  xb.l = (lane < w.nl) ? (io.lbar ? io.lbar[ol + lane] : x.l) : 0.0;
#pragma unroll
  for (int m = 0; m < NVR; m++) {
    const int k = lane + 32 * m;
    const bool in = k < w.nv;
    x.v[m] = (in && io.v) ? io.v[ov + k] : 0.0;
    x.y[m] = (in && io.y) ? io.y[ov + k] : 0.0;
    xb.v[m] = in ? (io.vbar ? io.vbar[ov + k] : x.v[m]) : 0.0;
    xb.y[m] = 0.0;
  }
  if (A.comp == FBSTAB_COMP_MARGIN) {
    w.publish(x);
    double az[NVR];
    w.Az(az);
#pragma unroll
    for (int m = 0; m < NVR; m++)
      if (lane + 32 * m < w.nv) io.dy[ov + lane + 32 * m] = w.br[m] - az[m];
  } else if (A.comp == FBSTAB_COMP_RESIDUAL) {
    EvalOut e = w.evaluate(x, xb, io.sigma, alpha, &ri);
    double sq[6];
    sq[0] = ri.z * ri.z;
    sq[1] = ri.l * ri.l;
    const double nzr = ri.z - io.sigma * (x.z - xb.z);
    const double nlr = ri.l - io.sigma * (x.l - xb.l);
    sq[3] = nzr * nzr;
    sq[4] = nlr * nlr;
    sq[2] = 0.0;
    sq[5] = 0.0;
#pragma unroll
    for (int m = 0; m < NVR; m++) {
      sq[2] = fma(ri.v[m], ri.v[m], sq[2]);
      const double n = pnr(x.y[m], x.v[m], alpha);
      sq[5] = fma(n, n, sq[5]);
    }
#pragma unroll
    for (int q = 0; q < 6; q++) sq[q] = sqrt(warp_sum(sq[q]));
    if (lane < w.nz) io.rz[oz + lane] = ri.z;
    if (lane < w.nl) io.rl[ol + lane] = ri.l;
#pragma unroll
    for (int m = 0; m < NVR; m++)
      if (lane + 32 * m < w.nv) io.rv[ov + lane + 32 * m] = ri.v[m];
    if (lane == 0 && io.norms) {
#pragma unroll
      for (int q = 0; q < 6; q++) io.norms[(size_t)inst * 8 + q] = sq[q];
      io.norms[(size_t)inst * 8 + 6] = e.Ei;
      io.norms[(size_t)inst * 8 + 7] = e.Eo;
    }
  } else if (A.comp == FBSTAB_COMP_NEWTON) {
    // the component API passes the right-hand side r; newton_step solves for -ri
    ri.z = (lane < w.nz) ? -io.rz[oz + lane] : 0.0;
    ri.l = (lane < w.nl) ? -io.rl[ol + lane] : 0.0;
#pragma unroll
    for (int m = 0; m < NVR; m++)
      ri.v[m] = (lane + 32 * m < w.nv) ? -io.rv[ov + lane + 32 * m] : 0.0;
    const bool okk = w.newton_step(x, xb, io.sigma, alpha, ri, &dx);
    if (lane < w.nz) io.dz[oz + lane] = dx.z;
    if (lane < w.nl) io.dl[ol + lane] = dx.l;
#pragma unroll
    for (int m = 0; m < NVR; m++) {
      const int k = lane + 32 * m;
      if (k < w.nv) {
        io.dv[ov + k] = dx.v[m];
        io.dy[ov + k] = dx.y[m];
        if (io.gamma) io.gamma[ov + k] = w.gamma[m];
        if (io.mus) io.mus[ov + k] = w.mus[m];
      }
    }
    if (lane == 0 && io.status)
      io.status[inst] = okk ? FBSTAB_STATUS_OK : FBSTAB_STATUS_FACTOR_FAILED;
  } else if (A.comp == FBSTAB_COMP_FEAS) {
    const int feas = w.feasibility(x, io.tol);
    if (lane == 0 && io.status) io.status[inst] = feas;
  }
}

// COMPONENT = false: the solver.  COMPONENT = true: one stage per instance
// (parity tests); a separate instantiation keeps the solver's code compact.
template <bool COMPONENT>
__global__ void __launch_bounds__(32 * kWarps, 1)
dense_small_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) double smem[];
  Warp w;
  w.lane = threadIdx.x & 31;
  w.sb = (unsigned)__cvta_generic_to_shared(smem) +
         (unsigned)(threadIdx.x >> 5) * (unsigned)(SLAB * sizeof(double));
  w.nz = a.nz;
  w.nl = a.nl;
  w.nv = a.nv;
  int inst = 0;
  if (w.lane == 0) inst = atomicAdd(a.counter, 1);
  inst = __shfl_sync(0xffffffffu, inst, 0);
  while (inst < a.batch) {
    w.load(a, inst);
    // reserve the next instance now and pull its data towards L2 meanwhile
    int next = 0;
    if (w.lane == 0) next = atomicAdd(a.counter, 1);
    next = __shfl_sync(0xffffffffu, next, 0);
    if (next < a.batch) w.prefetch(a, next);
    if (COMPONENT)
      run_component(w, a, inst);
    else
      solve_one(w, a, inst);
    inst = next;
  }
}

}  // namespace small

int DenseSmallWarpsPerCta() { return small::kWarps; }

int DenseSmallInit(DenseSmallPlan* p, int nz, int nl, int nv, int sm_count,
                   int* counter) {
  p->enabled = false;
  if (nz > small::NZ || nl > small::NL || nv > small::NV) return 0;
  p->nz = nz;
  p->nl = nl;
  p->nv = nv;
  p->counter = counter;
  p->smem = sizeof(double) * small::SLAB * small::kWarps;
  if (cudaFuncSetAttribute(small::dense_small_kernel<false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)p->smem) != cudaSuccess ||
      cudaFuncSetAttribute(small::dense_small_kernel<true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)p->smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;  // the generic kernel takes over
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &occ, small::dense_small_kernel<false>, 32 * small::kWarps, p->smem) !=
          cudaSuccess ||
      occ < 1) {
    cudaGetLastError();
    return 0;
  }
  p->grid = sm_count * occ;
  p->enabled = true;
  p->name = "dense-small-warp (smem-resident data, DMMA SYRK, register LDL')";
  return 0;
}

int DenseSmallLaunch(const DenseSmallPlan& p, int batch, const double* H,
                     const double* f, const double* G, const double* h,
                     const double* A, const double* b, double* z, double* l,
                     double* v, double* y, fbstab_out* out,
                     const fbstab_options& opts, int comp,
                     const fbstab_component_io* io, cudaStream_t stream) {
  small::Args a;
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
  const int ctas = (batch + small::kWarps - 1) / small::kWarps;
  const int grid = ctas < p.grid ? ctas : p.grid;
  if (comp < 0)
    small::dense_small_kernel<false><<<grid, 32 * small::kWarps, p.smem, stream>>>(a);
  else
    small::dense_small_kernel<true><<<grid, 32 * small::kWarps, p.smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

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
#include "lane_engine.cuh"
#include "mpc_lane.h"
#include "tma.cuh"

#ifdef PREFETCH_L1
#define PREFETCH_OP "prefetch.global.L1"
#else
#define PREFETCH_OP "prefetch.global.L2"
#endif
#ifndef PREFETCH_DIST
#define PREFETCH_DIST 1
#endif
#ifndef FBSTAB_LANE_WS_POLICY
#define FBSTAB_LANE_WS_POLICY 1
#endif
#ifndef FBSTAB_LANE_SDATA_SMEM
#define FBSTAB_LANE_SDATA_SMEM 1
#endif
#ifndef FBSTAB_LANE_SLOTS
#define FBSTAB_LANE_SLOTS 2
#endif

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
  // common-stage-data fast path: *mismatch == 0 <=> every instance of this launch
  // carries the stage data of instance 0, which `sdata` then holds stage-major
  const int* mismatch;
  const double* sdata;
  int sdata_smem;  // the launch carries (N+1)*DSZ doubles of dynamic shared memory
  int warps;       // warps that have a workspace (the last CTA may be partly idle)
};
constexpr int kLaneWarpsPerCta = 4;


// ---- register-resident small dense algebra ----------------------------------
// Same recursions as the team versions in mpc_riccati.cuh, with one change that
// matters for the latency of a lone instance: the Cholesky factors keep the
// RECIPROCAL of their diagonal in the diagonal slot, so that the ~40 dependent
// divisions per stage of the triangular solves become multiplications (an FP64
// division is a ~100-cycle dependent chain; a straggler instance runs its
// stages strictly one after the other).
//
// Which reciprocal matters for parity (profiles/r2_lane_diag_ab.txt, cfg 3a, 2,048 servo
// instances against the oracle; the oracle's own FMA on/off floor is 97.6 %):
//   FBSTAB_LANE_DIAG=0  rsqrt(x), multiply           255 k solves/s   95.7 % same trajectory
//                       (and one-sided: 69 of the 89 other instances end one Newton step EARLY)
//   FBSTAB_LANE_DIAG=1  sqrt(x), divide (the CPU reference's operation sequence, Eigen's
//                       unblocked llt_inplace)        122 k            97.9 %
//   FBSTAB_LANE_DIAG=2  1 / sqrt(x), both correctly rounded, multiply   (default)
//                                                     252 k            97.9 %
//   FBSTAB_LANE_DIAG=3  rsqrt(x) + one Newton step, multiply
//                                                     257 k            94.7 %
// The reciprocal of the ROUNDED square root -- the number the reference divides by -- keeps
// the factor consistent with the reference's; the better approximation of x^-1/2 does not.
#ifndef FBSTAB_LANE_DIAG
#define FBSTAB_LANE_DIAG 2
#endif
// x scaled by the inverse of a factor's diagonal entry d (the diagonal slot holds 1/d,
// or d itself under FBSTAB_LANE_DIAG=1)
__device__ __forceinline__ double by_diag(double x, double slot) {
#if FBSTAB_LANE_DIAG == 1
  return x / slot;
#else
  return x * slot;
#endif
}
// the diagonal slot of sqrt(x)
__device__ __forceinline__ double diag_slot(double x) {
#if FBSTAB_LANE_DIAG == 1
  return sqrt(x);
#elif FBSTAB_LANE_DIAG == 2
  return 1.0 / sqrt(x);
#elif FBSTAB_LANE_DIAG == 3
  const double r = rsqrt(x);  // r (1 + (1 - x r^2) / 2)
  const double e = fma(-(x * r), r, 1.0);
  return fma(0.5 * r, e, r);
#else
  return rsqrt(x);  // MUFU.RSQ64H + Newton: a third of the sqrt + division chain
#endif
}
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
    const double rx = diag_slot(x);
#pragma unroll
    for (int i = k + 1; i < M; i++) {
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < k; j++) a = fma(A[i][j], A[k][j], a);
      A[i][k] = by_diag(A[i][k] - a, rx);
    }
    A[k][k] = rx;  // diagonal slot
  }
  return ok;
}
// y = L^-1 x (x destroyed); L carries diagonal slots
template <int M>
__device__ __forceinline__ void trsv_l(const double (&L)[M][M], double (&x)[M], double (&y)[M]) {
#pragma unroll
  for (int j = 0; j < M; j++) {
    const double xj = by_diag(x[j], L[j][j]);
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
    const double xi = by_diag(x[i], L[i][i]);
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
    X[j] = by_diag(s, L[j][j]);
  }
}

// SHARED: every instance of the batch carries the same stage data (checked on
// the device before the launch, mpc_shared_detect): the data is read from one
// stage-major array common to all lanes (a broadcast load, L1 resident) and
// the per-lane transposed copy -- 39 % of the workspace of the servo problem,
// re-streamed from DRAM by every sweep -- does not exist.
template <int NX, int NU, int NC, bool SHARED = false>
struct Lane {
  static constexpr int NS = NX + NU;
  static constexpr int TX = NX * (NX + 1) / 2, TU = NU * (NU + 1) / 2;
  // factor block of a stage: lower L | lower M | AM | SM | P | lower SG
  static constexpr int oL = 0, oM = TX, oAM = 2 * TX, oSM = oAM + NX * NX,
                       oP = oSM + NU * NX, oSG = oP + NX * NU, FS = oSG + TU;
  // STAGE-MAJOR workspace: everything stage i needs is one contiguous block
  //   [ xk: z l v y | xi: z l v y | dx: z l v y | ri: z l v | gamma mu | factor ]
  // (element e of lane j at ws[e*32 + j]); a sweep touches consecutive blocks,
  // so the next block is prefetched into L2 while the current one is computed.
  static constexpr int V_Z = 0, V_L = NS, V_V = NS + NX, V_Y = NS + NX + NC,
                       VSZ = NS + NX + 2 * NC;
  static constexpr int O_XK = 0, O_XI = VSZ, O_DX = 2 * VSZ, O_RI = 3 * VSZ;
  static constexpr int R_Z = 0, R_L = NS, R_V = NS + NX, RSZ = NS + NX + NC;
  static constexpr int O_GM = O_RI + RSZ, O_FAC = O_GM + 2 * NC, O_DAT = O_FAC + FS;
  // ... | the stage's problem data, transposed once from the instance-major wire
  // format so that every later read is a coalesced, prefetchable access:
  //   Q | R | S | q | r | A | B | c | E | L | d
  static constexpr int D_Q = O_DAT, D_R = D_Q + NX * NX, D_S = D_R + NU * NU,
                       D_q = D_S + NU * NX, D_r = D_q + NX, D_A = D_r + NU,
                       D_B = D_A + NX * NX, D_c = D_B + NX * NU, D_E = D_c + NX,
                       D_L = D_E + NC * NX, D_d = D_L + NC * NU, SBF = D_d + NC;
  static constexpr int DSZ = SBF - O_DAT;            // data doubles per stage
  static constexpr int SB = SHARED ? O_DAT : SBF;    // per-lane block of a stage
  static constexpr int DS = SHARED ? 1 : 32;         // stride between data elements

  int N, nz, nl, nv;
  double* ws;  // this lane's column of the interleaved workspace
  const double *Q, *R, *S, *q, *r, *A, *B, *c, *E, *L, *d, *x0;
  const double* sdata;  // SHARED: [stage][DSZ] common stage data

  // element `o` of stage i's block
  // The streamed workspace bypasses L1 (ld.global.cg / st.global.cg), which then keeps
  // the common stage data and the local-memory spill slots (FBSTAB_LANE_WS_POLICY=0:
  // default caching, 2: evict-first hints -- A/B switches).
  __device__ __forceinline__ double ld(int i, int o) const {
#if FBSTAB_LANE_WS_POLICY == 1
    return __ldcg(ws + ((size_t)i * SB + o) * 32);
#elif FBSTAB_LANE_WS_POLICY == 2
    return __ldcs(ws + ((size_t)i * SB + o) * 32);
#else
    return ws[((size_t)i * SB + o) * 32];
#endif
  }
  // stores of a sweep that the whole warp executes are predicated on `on` (the lanes
  // the sweep is for)
  __device__ __forceinline__ void st(int i, int o, double v) const {
    if (!on) return;
#if FBSTAB_LANE_WS_POLICY == 1
    __stcg(ws + ((size_t)i * SB + o) * 32, v);
#elif FBSTAB_LANE_WS_POLICY == 2
    __stcs(ws + ((size_t)i * SB + o) * 32, v);
#else
    ws[((size_t)i * SB + o) * 32] = v;
#endif
  }

  // ---- stage-block ring --------------------------------------------------------
  // The four sweeps of a Newton round (evaluate, factor, forward and backward
  // substitution) read their stage blocks from a 2-slot ring in shared memory that the
  // TMA engine fills ONE STAGE AHEAD: a stage's block is contiguous in the
  // lane-interleaved workspace, so lane 0 issues one cp.async.bulk for the elements
  // [o0, o1) of the stage (plus up to two short runs, e.g. the next stage's
  // multipliers) and the slot's mbarrier counts the bytes.  The warp waits on shared
  // memory (~30 cycles) instead of on one L2 / DRAM round trip per batch of loads the
  // compiler could not hoist (59 % of all stall samples before,
  // profiles/r2_mpc_lane_ring.txt).  Stores go straight to global memory.
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int RING_E =
      cmax(cmax(3 * VSZ + 3 * NX, (O_FAC + FS - O_RI) + NX), (O_FAC + FS - (O_RI + R_V)) + NS + NX);
  static constexpr int SLOTS = FBSTAB_LANE_SLOTS;  // ring depth: SLOTS - 1 stages in flight
  double* ring;        // this lane's column of slot 0 (element e of slot s: ring[(s*RING_E+e)*32])
  unsigned bar0;       // shared address of the SLOTS slot mbarriers
  unsigned par;        // bit s: the parity the next wait on slot s expects
  bool on;             // this lane takes part in the current sweep (stores enabled)
  static constexpr bool enabled = true;  // every lane of a warp owns instances

  // sweep start: the stores of earlier sweeps become visible to the TMA engine
  __device__ __forceinline__ void ring_begin() const {
    tma::fence_proxy_async_all();
    __syncwarp();
  }
  // stage i's elements [o0, o1) and up to three runs of xn elements of stage ix, starting
  // at x0o, x1o, x2o (a negative offset: no run) -> slot s, back to back
  __device__ __forceinline__ void ring_issue(int s, int i, int o0, int o1, int ix = -1,
                                             int xn = 0, int x0o = -1, int x1o = -1,
                                             int x2o = -1) const {
    if ((threadIdx.x & 31) == 0) {
      const bool hx = ix >= 0 && ix <= N && xn > 0;
      const int xo[3] = {x0o, x1o, x2o};
      int runs = 0;
#pragma unroll
      for (int m = 0; m < 3; m++) runs += (hx && xo[m] >= 0) ? 1 : 0;
      const unsigned bar = bar0 + 8u * s;
      unsigned dst = tma::smem_addr(ring) + (unsigned)s * RING_E * 256u;
      const char* base = (const char*)ws;
      tma::mbar_arrive_expect_tx(bar, (unsigned)(o1 - o0 + runs * xn) * 256u);
      tma::bulk_g2s(dst, base + ((size_t)i * SB + o0) * 256, (unsigned)(o1 - o0) * 256u, bar);
      dst += (unsigned)(o1 - o0) * 256u;
#pragma unroll
      for (int m = 0; m < 3; m++)
        if (hx && xo[m] >= 0) {
          tma::bulk_g2s(dst, base + ((size_t)ix * SB + xo[m]) * 256, (unsigned)xn * 256u, bar);
          dst += (unsigned)xn * 256u;
        }
    }
  }
  // waits for slot s; returns this lane's column of it
  __device__ __forceinline__ const double* ring_wait(int s) {
    tma::mbar_wait(bar0 + 8u * s, (par >> s) & 1u);
    par ^= 1u << s;
    return ring + (size_t)s * RING_E * 32;
  }

  // address of element `o` of stage i's block (stride 32 doubles between elements)
  __device__ __forceinline__ const double* dat(int i, int o) const {
    if (SHARED) return sdata + (size_t)i * DSZ + (o - O_DAT);
    return ws + ((size_t)i * SB + o) * 32;
  }
  // L2 prefetch of elements [o0, o1) of stage i's block: the warp's 32 lanes
  // cover the (o1-o0)*256 bytes line by line
  __device__ __forceinline__ void prefetch(int i, int o0, int o1) const {
    if (i < 0 || i > N) return;
    if (o1 > SB) o1 = SB;  // SHARED: there is no per-lane data block
    if (o0 >= o1) return;
    const int lane = threadIdx.x & 31;
    const char* base = (const char*)(ws - lane) + ((size_t)i * SB + o0) * 256;
    const int lines = (o1 - o0) * 2;
    for (int m = lane; m < lines; m += 32)
      asm volatile(PREFETCH_OP " [%0];" ::"l"(base + (size_t)m * 128));
  }

  __device__ void bind(const LaneArgs& a, int inst) {
    // SHARED: every instance carries instance 0's stage data -- and the caller may have
    // passed exactly ONE copy (fbstab_mpc_batch_solve_shared); only x0 is per instance
    const size_t i = SHARED ? 0 : (size_t)inst, K = N + 1;
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
    x0 = a.data.x0 + (size_t)inst * NX;
  }

  // packed lower triangles <-> register matrices (t = FS-sized register copy)
  __device__ __forceinline__ static void unpack_x(const double* t, double (&M)[NX][NX]) {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NX; cc++)
#pragma unroll
      for (int rr = cc; rr < NX; rr++) M[rr][cc] = t[k++];
  }
  __device__ __forceinline__ static void unpack_u(const double* t, double (&M)[NU][NU]) {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NU; cc++)
#pragma unroll
      for (int rr = cc; rr < NU; rr++) M[rr][cc] = t[k++];
  }
  __device__ __forceinline__ void store_lower(int i, int o, const double (&M)[NX][NX]) const {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NX; cc++)
#pragma unroll
      for (int rr = cc; rr < NX; rr++) st(i, o + k++, M[rr][cc]);
  }
  __device__ __forceinline__ void store_lower_u(int i, int o, const double (&M)[NU][NU]) const {
    int k = 0;
#pragma unroll
    for (int cc = 0; cc < NU; cc++)
#pragma unroll
      for (int rr = cc; rr < NU; rr++) st(i, o + k++, M[rr][cc]);
  }

  // (E(i) x + L(i) u)[k], mpc_data.cc:66-105
  __device__ __forceinline__ double Az_entry(const double (&z)[NS], int i, int k) const {
    const double* Em = dat(i, D_E);
    const double* Lm = dat(i, D_L);
    double s = 0.0, s2 = 0.0;
#pragma unroll
    for (int cc = 0; cc < NX; cc++) s = fma(Em[(k + cc * NC) * DS], z[cc], s);
#pragma unroll
    for (int cc = 0; cc < NU; cc++) s2 = fma(Lm[(k + cc * NC) * DS], z[NX + cc], s2);
    return s + s2;
  }

  // CopyIntoVariable + InitializeConstraintMargin + xi = xk; returns the forcing norm
  __device__ double init(const double* z0, const double* l0, const double* v0) {
    double fn = 0.0;
    for (int i = 0; i <= N; i++) {
      // stage data: instance-major wire format -> this lane's workspace column
      if (!SHARED) {
        const double* src[11] = {Q + (size_t)i * NX * NX, R + (size_t)i * NU * NU,
                                 S + (size_t)i * NU * NX, q + (size_t)i * NX,
                                 r + (size_t)i * NU,      A + (size_t)i * NX * NX,
                                 B + (size_t)i * NX * NU, c + (size_t)i * NX,
                                 E + (size_t)i * NC * NX, L + (size_t)i * NC * NU,
                                 d + (size_t)i * NC};
        const int off[11] = {D_Q, D_R, D_S, D_q, D_r, D_A, D_B, D_c, D_E, D_L, D_d};
        const int cnt[11] = {NX * NX, NU * NU, NU * NX, NX, NU, NX * NX, NX * NU, NX,
                             NC * NX, NC * NU, NC};
#pragma unroll
        for (int a_ = 0; a_ < 11; a_++) {
          const bool stage_only = (a_ == 5 || a_ == 6 || a_ == 7);  // A, B, c: N entries
          if (!stage_only || i < N) {
#pragma unroll
            for (int k = 0; k < cnt[a_]; k++) st(i, off[a_] + k, __ldg(src[a_] + k));
          }
        }
      }
      double z[NS], lv[NX], vv[NC], yv[NC];
#pragma unroll
      for (int k = 0; k < NS; k++) z[k] = z0[i * NS + k];
#pragma unroll
      for (int k = 0; k < NX; k++) {
        lv[k] = l0[i * NX + k];
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
        vv[k] = v0[i * NC + k];
        const double dv = __ldg(d + i * NC + k);
        fn = fma(dv, dv, fn);
        yv[k] = -dv - Az_entry(z, i, k);
      }
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int o = b ? O_XI : O_XK;
#pragma unroll
        for (int k = 0; k < NS; k++) st(i, o + V_Z + k, z[k]);
#pragma unroll
        for (int k = 0; k < NX; k++) st(i, o + V_L + k, lv[k]);
#pragma unroll
        for (int k = 0; k < NC; k++) {
          st(i, o + V_V + k, vv[k]);
          st(i, o + V_Y + k, yv[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NX; k++) fn = fma(__ldg(x0 + k), __ldg(x0 + k), fn);
    return sqrt(fn);
  }

  // Fused residual evaluation at x = base (+ t dx when trial): inner residual
  // wrt xbar = xk (or wrt x itself when self_bar) -> ri, both norms.  Every
  // load of a stage is issued before its first store (the compiler cannot
  // move a load across a store to the same array, and each exposed load is a
  // DRAM round trip).
  __device__ EvalOut evaluate(int base, bool trial, double t, bool self_bar, double sigma,
                              double alpha) {
    const double sg = self_bar ? 0.0 : sigma;
    double s[6] = {0, 0, 0, 0, 0, 0};
    double zp[NS], lc[NX], ln[NX];  // z(i-1), l(i), l(i+1)
#pragma unroll
    for (int k = 0; k < NX; k++) lc[k] = ln[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NS; k++) zp[k] = 0.0;
    // ring: [ xk | xi | dx ] of the stage, then l of the next stage in xk, xi and dx
    constexpr int E0 = 3 * VSZ;
    const int lsel = (base == O_XK) ? 0 : NX;
    ring_begin();
    for (int j = 0; j < SLOTS - 1 && j <= N; j++)
      ring_issue(j, j, 0, E0, j + 1, NX, O_XK + V_L, O_XI + V_L, O_DX + V_L);
    for (int i = 0; i <= N; i++) {
      prefetch(i + PREFETCH_DIST, O_DAT, SBF);
      __syncwarp();
      {
        const int j = i + SLOTS - 1;
        if (j <= N) ring_issue(j % SLOTS, j, 0, E0, j + 1, NX, O_XK + V_L, O_XI + V_L, O_DX + V_L);
      }
      const double* sl = ring_wait(i % SLOTS);
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < NX; k++) {
          const double a0 = sl[(base + V_L + k) * 32], b0 = sl[(O_DX + V_L + k) * 32];
          lc[k] = trial ? fma(t, b0, a0) : a0;
        }
      }
      double xb[VSZ], xd[VSZ], zk[NS], lk[NX], vk[NC], la[NX], lb[NX];
#pragma unroll
      for (int k = 0; k < VSZ; k++) {
        xb[k] = sl[(base + k) * 32];
        xd[k] = sl[(O_DX + k) * 32];
      }
#pragma unroll
      for (int k = 0; k < NS; k++) zk[k] = sl[(O_XK + V_Z + k) * 32];
#pragma unroll
      for (int k = 0; k < NX; k++) lk[k] = sl[(O_XK + V_L + k) * 32];
#pragma unroll
      for (int k = 0; k < NC; k++) vk[k] = sl[(O_XK + V_V + k) * 32];
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++) {
          la[k] = sl[(E0 + lsel + k) * 32];
          lb[k] = sl[(E0 + 2 * NX + k) * 32];
        }
      }
      double z[NS], v[NC], y[NC];
#pragma unroll
      for (int k = 0; k < NS; k++) z[k] = trial ? fma(t, xd[V_Z + k], xb[V_Z + k]) : xb[V_Z + k];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        v[k] = trial ? fma(t, xd[V_V + k], xb[V_V + k]) : xb[V_V + k];
        // y-aware axpy, full_variable.cc:55-65
        const double y1 = fma(t, xd[V_Y + k], xb[V_Y + k]);
        const double y2 = fma(-t, -dat(i, D_d)[k * DS], y1);
        y[k] = trial ? y2 : xb[V_Y + k];
      }
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++) ln[k] = trial ? fma(t, lb[k], la[k]) : la[k];
      }
      const double* Qm = dat(i, D_Q);
      const double* Rm = dat(i, D_R);
      const double* Sm = dat(i, D_S);
      const double* Em = dat(i, D_E);
      const double* Lm = dat(i, D_L);
      const double* Am = dat(i, D_A);
      const double* Bm = dat(i, D_B);
      double out[RSZ];
      // z block: tz = ((f + Hz) + G'l) + A'v
#pragma unroll
      for (int rr = 0; rr < NS; rr++) {
        double s1 = 0.0, s2 = 0.0, tz;
        if (rr < NX) {
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s1 = fma(Qm[(rr + cc * NX) * DS], z[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s2 = fma(Sm[(cc + rr * NU) * DS], z[NX + cc], s2);
          tz = dat(i, D_q)[rr * DS] + (s1 + s2);
          tz += -lc[rr];
          if (i < N) {
            double sa = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) sa = fma(Am[(cc + rr * NX) * DS], ln[cc], sa);
            tz += sa;
          }
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Em[(k + rr * NC) * DS], v[k], sv);
          tz += sv;
        } else {
          const int ru = rr - NX;
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s1 = fma(Sm[(ru + cc * NU) * DS], z[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s2 = fma(Rm[(ru + cc * NU) * DS], z[NX + cc], s2);
          tz = dat(i, D_r)[ru * DS] + (s1 + s2);
          if (i < N) {
            double sa = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) sa = fma(Bm[(cc + ru * NX) * DS], ln[cc], sa);
            tz += sa;
          }
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Lm[(k + ru * NC) * DS], v[k], sv);
          tz += sv;
        }
        s[3] = fma(tz, tz, s[3]);
        const double rr_ = tz + sg * (z[rr] - zk[rr]);
        out[R_Z + rr] = rr_;
        s[0] = fma(rr_, rr_, s[0]);
      }
      // l block: tl = h - Gz
#pragma unroll
      for (int rr = 0; rr < NX; rr++) {
        double tl;
        if (i == 0) {
          tl = -__ldg(x0 + rr) + z[rr];
        } else {
          const double* Ap = dat(i - 1, D_A);
          const double* Bp = dat(i - 1, D_B);
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s1 = fma(Ap[(rr + cc * NX) * DS], zp[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s2 = fma(Bp[(rr + cc * NX) * DS], zp[NX + cc], s2);
          tl = (-dat(i - 1, D_c)[rr * DS] - (s1 + s2)) + z[rr];
        }
        s[4] = fma(tl, tl, s[4]);
        const double rl = tl + sg * (lc[rr] - lk[rr]);
        out[R_L + rr] = rl;
        s[1] = fma(rl, rl, s[1]);
      }
      // v block
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const double ys = y[k] + sg * (v[k] - vk[k]);
        const double rv = pfb(ys, v[k], alpha);
        out[R_V + k] = rv;
        s[2] = fma(rv, rv, s[2]);
        const double nn = pnr(y[k], v[k], alpha);
        s[5] = fma(nn, nn, s[5]);
      }
#pragma unroll
      for (int k = 0; k < RSZ; k++) st(i, O_RI + k, out[k]);
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
      prefetch(i + PREFETCH_DIST, O_XI, O_RI);
      prefetch(i + PREFETCH_DIST, D_d, SBF);
      double xb[VSZ], xd[VSZ];
#pragma unroll
      for (int k = 0; k < VSZ; k++) {
        xb[k] = ld(i, O_XI + k);
        xd[k] = ld(i, O_DX + k);
      }
#pragma unroll
      for (int k = 0; k < V_Y; k++) xb[k] = fma(t, xd[k], xb[k]);
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const double y1 = fma(t, xd[V_Y + k], xb[V_Y + k]);
        xb[V_Y + k] = fma(-t, -dat(i, D_d)[k * DS], y1);
      }
#pragma unroll
      for (int k = 0; k < VSZ; k++) st(i, O_XI + k, xb[k]);
    }
  }

  // RiccatiLinearSolver::Initialize, riccati_linear_solver.cc:77-210.
  // with_commit: the accepted step xi <- xi + t dx is applied in the same sweep
  // (saves one pass over the horizon in the common accept-then-Newton round).
  // any_commit: some lane of the warp commits (warp-uniform; selects the ring range)
  __device__ bool factor(double sigma, double alpha, bool with_commit, double t,
                         bool any_commit) {
    bool ok = true;
    double Lc[NX][NX];
#if FBSTAB_LANE_DIAG == 1
    const double rs = sqrt(sigma);  // diagonal slot of L(0) = sqrt(sigma) I
#else
    const double rs = 1.0 / sqrt(sigma);  // reciprocal diagonal of L(0) = sqrt(sigma) I
#endif
#pragma unroll
    for (int a_ = 0; a_ < NX; a_++)
#pragma unroll
      for (int b_ = 0; b_ < NX; b_++) Lc[a_][b_] = (a_ == b_) ? rs : 0.0;
    // ring: xi.v, xi.y (with a commit in the warp: all of xi and dx), then xk.v
    const int o0 = any_commit ? O_XI : O_XI + V_V;
    const int o1 = any_commit ? O_RI : O_XI + VSZ;
    const int ox = o1 - o0;
    ring_begin();
    for (int j = 0; j < SLOTS - 1 && j <= N; j++) ring_issue(j, j, o0, o1, j, NC, O_XK + V_V);
    for (int i = 0; i <= N; i++) {
      prefetch(i + PREFETCH_DIST, O_DAT, SBF);
      __syncwarp();
      {
        const int j = i + SLOTS - 1;
        if (j <= N) ring_issue(j % SLOTS, j, o0, o1, j, NC, O_XK + V_V);
      }
      const double* sl = ring_wait(i % SLOTS);
      double yv[NC], vv[NC], vk[NC];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        yv[k] = sl[(O_XI + V_Y + k - o0) * 32];
        vv[k] = sl[(O_XI + V_V + k - o0) * 32];
        vk[k] = sl[(ox + k) * 32];
      }
      if (any_commit) {
        double xz[V_V], dz[V_V], dv[NC], dy[NC], dd[NC];
#pragma unroll
        for (int k = 0; k < V_V; k++) {
          xz[k] = sl[(O_XI + k - O_XI) * 32];
          dz[k] = sl[(O_DX + k - O_XI) * 32];
        }
#pragma unroll
        for (int k = 0; k < NC; k++) {
          dv[k] = sl[(O_DX + V_V + k - O_XI) * 32];
          dy[k] = sl[(O_DX + V_Y + k - O_XI) * 32];
          dd[k] = dat(i, D_d)[k * DS];
        }
        const bool keep = on;
        on = keep && with_commit;
#pragma unroll
        for (int k = 0; k < V_V; k++) st(i, O_XI + k, fma(t, dz[k], xz[k]));
#pragma unroll
        for (int k = 0; k < NC; k++) {
          const double v2 = fma(t, dv[k], vv[k]);
          const double y2 = fma(-t, -dd[k], fma(t, dy[k], yv[k]));
          vv[k] = with_commit ? v2 : vv[k];
          yv[k] = with_commit ? y2 : yv[k];
          st(i, O_XI + V_V + k, vv[k]);
          st(i, O_XI + V_Y + k, yv[k]);
        }
        on = keep;
      }
      store_lower(i, O_FAC + oL, Lc);
      double Gam[NC];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const double ys = yv[k] + sigma * (vv[k] - vk[k]);
        double ga, mu;
        pfb_barrier(ys, vv[k], alpha, sigma, &ga, &mu);
        st(i, O_GM + k, ga);
        st(i, O_GM + NC + k, mu);
        Gam[k] = div_nr(ga, mu);
      }
      const double* Qm = dat(i, D_Q);
      const double* Rm = dat(i, D_R);
      const double* Sm = dat(i, D_S);
      const double* Em = dat(i, D_E);
      const double* Lm = dat(i, D_L);
      double Ee[NC][NX], Le[NC][NU];
#pragma unroll
      for (int k = 0; k < NC; k++) {
#pragma unroll
        for (int cc = 0; cc < NX; cc++) Ee[k][cc] = Em[(k + cc * NC) * DS];
#pragma unroll
        for (int cc = 0; cc < NU; cc++) Le[k][cc] = Lm[(k + cc * NC) * DS];
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
          w[j] = by_diag(w[j], Lc[j][j]);
          const double wj = w[j];
#pragma unroll
          for (int k = j + 1; k < NX; k++) w[k] = fma(-Lc[k][j], wj, w[k]);
        }
#pragma unroll
        for (int k = NX - 1; k >= 0; k--) {
          double sv = w[k];
#pragma unroll
          for (int j = k + 1; j < NX; j++) sv = fma(-Lc[j][k], w[j], sv);
          w[k] = by_diag(sv, Lc[k][k]);
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
          const double qt = (Qm[(rr + cc * NX) * DS] + (rr == cc ? sigma : 0.0)) + sv;
          Mm[rr][cc] = qt + Mm[rr][cc];
        }
      ok = chol<NX>(Mm) && ok;
      store_lower(i, O_FAC + oM, Mm);
      // R~, S~
      double Rt[NU][NU], St[NU][NX];
#pragma unroll
      for (int cc = 0; cc < NU; cc++)
#pragma unroll
        for (int rr = cc; rr < NU; rr++) {
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Le[k][rr], Gam[k] * Le[k][cc], sv);
          Rt[rr][cc] = (Rm[(rr + cc * NU) * DS] + (rr == cc ? sigma : 0.0)) + sv;
        }
#pragma unroll
      for (int cc = 0; cc < NX; cc++)
#pragma unroll
        for (int rr = 0; rr < NU; rr++) {
          double sv = 0.0;
#pragma unroll
          for (int k = 0; k < NC; k++) sv = fma(Le[k][rr], Gam[k] * Ee[k][cc], sv);
          St[rr][cc] = Sm[(rr + cc * NU) * DS] + sv;
        }
      // AM = A M^-T, SM = S~ M^-T (:149-161)
      double AM[NX][NX], SM[NU][NX];
      if (i < N) {
        const double* Am = dat(i, D_A);
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double src[NX];
#pragma unroll
          for (int cc = 0; cc < NX; cc++) src[cc] = Am[(rr + cc * NX) * DS];
          row_trsm_lt<NX>(Mm, src, AM[rr]);
        }
#pragma unroll
        for (int rr = 0; rr < NX; rr++)
#pragma unroll
          for (int cc = 0; cc < NX; cc++) st(i, O_FAC + oAM + rr + cc * NX, AM[rr][cc]);
      }
#pragma unroll
      for (int rr = 0; rr < NU; rr++) row_trsm_lt<NX>(Mm, St[rr], SM[rr]);
#pragma unroll
      for (int rr = 0; rr < NU; rr++)
#pragma unroll
        for (int cc = 0; cc < NX; cc++) st(i, O_FAC + oSM + rr + cc * NU, SM[rr][cc]);
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
      store_lower_u(i, O_FAC + oSG, SG);
      if (i == N) break;
      // P = (AM SM' - B) SG^-T (:170-175)
      double P[NX][NU];
      {
        const double* Bm = dat(i, D_B);
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double src[NU];
#pragma unroll
          for (int j = 0; j < NU; j++) {
            double sv = 0.0;
#pragma unroll
            for (int k = 0; k < NX; k++) sv = fma(AM[rr][k], SM[j][k], sv);
            src[j] = sv - Bm[(rr + j * NX) * DS];
          }
          row_trsm_lt<NU>(SG, src, P[rr]);
#pragma unroll
          for (int j = 0; j < NU; j++) st(i, O_FAC + oP + rr + j * NX, P[rr][j]);
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

  // dv = (rv + gamma .* A dz) ./ mus ; dy = b - A dz for one stage (:331-341);
  // rv, ga, mu were loaded before the stage's first store
  __device__ __forceinline__ void finish_stage(int i, const double (&dz)[NS],
                                               const double (&rv)[NC], const double (&ga)[NC],
                                               const double (&mu)[NC]) {
#pragma unroll
    for (int k = 0; k < NC; k++) {
      const double sv = Az_entry(dz, i, k);
      st(i, O_DX + V_V + k, div_nr((-rv[k]) + ga[k] * sv, mu[k]));
      st(i, O_DX + V_Y + k, (-sv) + (-dat(i, D_d)[k * DS]));
    }
  }

  // RiccatiLinearSolver::Solve on r = -(ri), riccati_linear_solver.cc:212-344.
  // Storage reuse as in mpc_riccati.cuh: theta(i) -> dx.l(i), M^-1 h -> dx.z x(i),
  // SG^-1(.) -> dx.z u(i) until the backward sweep writes the step there.
  __device__ void solve() {
    double th[NX];  // theta(i)
    double lp[NX];  // dl(i+1) in the backward sweep
    // ring: [ ri | gamma mu | factor ] of the stage, then ri.l of the next stage
    constexpr int F0 = O_RI, F1 = O_FAC + FS, FX = F1 - F0;
    ring_begin();
    for (int j = 0; j < SLOTS - 1 && j <= N; j++) ring_issue(j, j, F0, F1, j + 1, NX, O_RI + R_L);
    for (int i = 0; i <= N; i++) {
      prefetch(i + PREFETCH_DIST, O_DAT, SBF);
      __syncwarp();
      {
        const int j = i + SLOTS - 1;
        if (j <= N) ring_issue(j % SLOTS, j, F0, F1, j + 1, NX, O_RI + R_L);
      }
      const double* sl = ring_wait(i % SLOTS);
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < NX; k++) th[k] = sl[(O_RI + R_L + k - F0) * 32];  // r2(0) = rl(0)
      }
      double rr_[RSZ], mu[NC], ga[NC], fa[FS], rln[NX];
#pragma unroll
      for (int k = 0; k < RSZ; k++) rr_[k] = sl[(O_RI + k - F0) * 32];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        ga[k] = sl[(O_GM + k - F0) * 32];
        mu[k] = sl[(O_GM + NC + k - F0) * 32];
      }
#pragma unroll
      for (int k = 0; k < FS; k++) fa[k] = sl[(O_FAC + k - F0) * 32];
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++) rln[k] = sl[(FX + k) * 32];
      }
      // r3 = rv ./ mus, r1 = r.z - A' r3 for this stage (:222-225)
      double tv[NC], r1[NS], rvv[NC];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        rvv[k] = rr_[R_V + k];
        tv[k] = div_nr(-rvv[k], mu[k]);
      }
      {
        const double* Em = dat(i, D_E);
        const double* Lm = dat(i, D_L);
#pragma unroll
        for (int rr = 0; rr < NS; rr++) {
          double sv = 0.0;
          if (rr < NX) {
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(Em[(k + rr * NC) * DS], tv[k], sv);
          } else {
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(Lm[(k + (rr - NX) * NC) * DS], tv[k], sv);
          }
          r1[rr] = (-rr_[R_Z + rr]) - sv;
        }
      }
      double Lc[NX][NX], Mm[NX][NX], SG[NU][NU], SM[NU][NX];
      unpack_x(fa + oL, Lc);
      unpack_x(fa + oM, Mm);
      unpack_u(fa + oSG, SG);
#pragma unroll
      for (int rr = 0; rr < NU; rr++)
#pragma unroll
        for (int cc = 0; cc < NX; cc++) SM[rr][cc] = fa[oSM + rr + cc * NU];
      // h(i) = (L L')^-1 theta(i) - r1x(i)
      double sa[NX], sb[NX], sc[NX], tx[NX], thi[NX];
#pragma unroll
      for (int k = 0; k < NX; k++) {
        sa[k] = th[k];
        thi[k] = th[k];
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
        // theta(i+1) = (P tu + AM tx) + r2(i+1)
#pragma unroll
        for (int k = 0; k < NX; k++) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int cc = 0; cc < NU; cc++) s1 = fma(fa[oP + k + cc * NX], tu[cc], s1);
#pragma unroll
          for (int cc = 0; cc < NX; cc++) s2 = fma(fa[oAM + k + cc * NX], tx[cc], s2);
          th[k] = (s1 + s2) + rln[k];
        }
#pragma unroll
        for (int k = 0; k < NX; k++) {
          st(i, O_DX + V_L + k, thi[k]);
          st(i, O_DX + V_Z + k, tx[k]);
        }
#pragma unroll
        for (int k = 0; k < NU; k++) st(i, O_DX + V_Z + NX + k, tu[k]);
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
          sa[k] = xN[k] + thi[k];
        }
        trsv_l<NX>(Lc, sa, sb);
        trsv_lt<NX>(Lc, sb, sc);
        double zN[NS];
#pragma unroll
        for (int k = 0; k < NX; k++) {
          lp[k] = -sc[k];
          zN[k] = xN[k];
        }
#pragma unroll
        for (int k = 0; k < NU; k++) zN[NX + k] = uN[k];
#pragma unroll
        for (int k = 0; k < NX; k++) st(i, O_DX + V_L + k, lp[k]);
#pragma unroll
        for (int k = 0; k < NS; k++) st(i, O_DX + V_Z + k, zN[k]);
        finish_stage(i, zN, rvv, ga, mu);
      }
    }
    // backward recursion :297-327
    // ring: [ ri.v | gamma mu | factor ] of the stage, then its dx.z, dx.l (the forward
    // sweep's M^-1 h, SG^-1(.) and theta)
    constexpr int B0 = O_RI + R_V, B1 = O_FAC + FS, BX = B1 - B0;
    ring_begin();
    for (int j = 0; j < SLOTS - 1 && j < N; j++)
      ring_issue(j, N - 1 - j, B0, B1, N - 1 - j, NS + NX, O_DX + V_Z);
    for (int i = N - 1, kk = 0; i >= 0; i--, kk++) {
      prefetch(i - PREFETCH_DIST, O_DAT, SBF);
      __syncwarp();
      {
        const int j = kk + SLOTS - 1;
        if (j < N) ring_issue(j % SLOTS, N - 1 - j, B0, B1, N - 1 - j, NS + NX, O_DX + V_Z);
      }
      const double* sl = ring_wait(kk % SLOTS);
      double fa[FS], dxz[NS], thi[NX], rvv[NC], ga[NC], mu[NC];
#pragma unroll
      for (int k = 0; k < FS; k++) fa[k] = sl[(O_FAC + k - B0) * 32];
#pragma unroll
      for (int k = 0; k < NS; k++) dxz[k] = sl[(BX + V_Z + k) * 32];
#pragma unroll
      for (int k = 0; k < NX; k++) thi[k] = sl[(BX + V_L + k) * 32];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        rvv[k] = sl[(O_RI + R_V + k - B0) * 32];
        ga[k] = sl[(O_GM + k - B0) * 32];
        mu[k] = sl[(O_GM + NC + k - B0) * 32];
      }
      double Lc[NX][NX], Mm[NX][NX], SG[NU][NU];
      unpack_x(fa + oL, Lc);
      unpack_x(fa + oM, Mm);
      unpack_u(fa + oSG, SG);
      double ua[NU], ui[NU], sa[NX], sb[NX], sc[NX], xi_[NX];
#pragma unroll
      for (int k = 0; k < NU; k++) {
        double sv = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; rr++) sv = fma(fa[oP + rr + k * NX], lp[rr], sv);
        ua[k] = dxz[NX + k] + sv;
      }
      trsv_lt<NU>(SG, ua, ui);
#pragma unroll
      for (int k = 0; k < NX; k++) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int rr = 0; rr < NU; rr++) s1 = fma(fa[oSM + rr + k * NU], ui[rr], s1);
#pragma unroll
        for (int rr = 0; rr < NX; rr++) s2 = fma(fa[oAM + rr + k * NX], lp[rr], s2);
        sa[k] = (dxz[k] + s1) + s2;
      }
      trsv_lt<NX>(Mm, sa, sb);
#pragma unroll
      for (int k = 0; k < NX; k++) {
        xi_[k] = -sb[k];
        sa[k] = thi[k] + xi_[k];
      }
      trsv_l<NX>(Lc, sa, sb);
      trsv_lt<NX>(Lc, sb, sc);
      double zi[NS];
#pragma unroll
      for (int k = 0; k < NX; k++) {
        lp[k] = -sc[k];
        zi[k] = xi_[k];
      }
#pragma unroll
      for (int k = 0; k < NU; k++) zi[NX + k] = ui[k];
#pragma unroll
      for (int k = 0; k < NX; k++) st(i, O_DX + V_L + k, lp[k]);
#pragma unroll
      for (int k = 0; k < NS; k++) st(i, O_DX + V_Z + k, zi[k]);
      finish_stage(i, zi, rvv, ga, mu);
    }
  }

  // End of a proximal subproblem in ONE sweep over the ring (impl:301, 202-216):
  // ProjectDuals on xi; then, for the lanes with do_diff (not at the Newton cap),
  // dx = xi - xk (y-aware), its norm, FullFeasibility::CheckFeasibility on it
  // (full_feasibility.cc:25-88; status in *feas) and xk <- xi.  (The copy is
  // unconditional for those lanes: after an infeasibility exit the result is dx and
  // xk is not read again.)
  // (with_commit: the lane's accepted step xi <- xi + t dx is applied in the same sweep,
  // with the fused multiply-adds of commit(); any_commit: some lane of the warp does.)
  __device__ double prox_end(bool do_diff, bool with_commit, double t, bool any_commit,
                             double tol, bool check, int* feas) {
    // ring: [ xk | xi ] of the stage (with a commit in the warp: | dx), then l of the
    // next stage in xk, xi (and dx)
    constexpr int P0 = O_XK;
    const int P1 = any_commit ? O_RI : O_XI + VSZ, PX = P1 - P0;
    const bool lanes = on;
    double s[3] = {0, 0, 0};
    double mx0 = -INFINITY, mx1 = 0, mx2 = 0, mx3 = 0, mp0 = 0, mp1 = 0, mp2 = 0;
    double sm0 = 0, sm1 = 0;
    double zp[NS], lc[NX], ln[NX];
#pragma unroll
    for (int k = 0; k < NS; k++) zp[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NX; k++) lc[k] = ln[k] = 0.0;
    const int xdl = any_commit ? O_DX + V_L : -1;
    ring_begin();
    for (int j = 0; j < SLOTS - 1 && j <= N; j++)
      ring_issue(j, j, P0, P1, j + 1, NX, O_XK + V_L, O_XI + V_L, xdl);
    for (int i = 0; i <= N; i++) {
      if (check) prefetch(i + PREFETCH_DIST, O_DAT, SBF);
      __syncwarp();
      {
        const int j = i + SLOTS - 1;
        if (j <= N) ring_issue(j % SLOTS, j, P0, P1, j + 1, NX, O_XK + V_L, O_XI + V_L, xdl);
      }
      const double* sl = ring_wait(i % SLOTS);
      double xa[VSZ], xb[VSZ], la[NX], lb[NX];
#pragma unroll
      for (int k = 0; k < VSZ; k++) {
        xa[k] = sl[(O_XI + k - P0) * 32];
        xb[k] = sl[(O_XK + k - P0) * 32];
      }
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++) {
          la[k] = sl[(PX + NX + k) * 32];
          lb[k] = sl[(PX + k) * 32];
        }
      }
      if (any_commit) {
        // xi <- xi + t dx (commit(): y-aware, full_variable.cc:55-65)
        on = lanes && with_commit;
#pragma unroll
        for (int k = 0; k < V_Y; k++) {
          const double c2 = fma(t, sl[(O_DX + k - P0) * 32], xa[k]);
          xa[k] = with_commit ? c2 : xa[k];
        }
#pragma unroll
        for (int k = 0; k < NC; k++) {
          const double y1 = fma(t, sl[(O_DX + V_Y + k - P0) * 32], xa[V_Y + k]);
          const double y2 = fma(-t, -dat(i, D_d)[k * DS], y1);
          xa[V_Y + k] = with_commit ? y2 : xa[V_Y + k];
        }
        if (i < N) {
#pragma unroll
          for (int k = 0; k < NX; k++) {
            const double c2 = fma(t, sl[(PX + 2 * NX + k) * 32], la[k]);
            la[k] = with_commit ? c2 : la[k];
          }
        }
        // (v is stored below, after the projection)
#pragma unroll
        for (int k = 0; k < V_V; k++) st(i, O_XI + k, xa[k]);
#pragma unroll
        for (int k = 0; k < NC; k++) st(i, O_XI + V_Y + k, xa[V_Y + k]);
      }
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < NX; k++) lc[k] = xa[V_L + k] + (-1.0) * xb[V_L + k];
      }
      // ProjectDuals (full_variable.cc:75)
      on = lanes;
#pragma unroll
      for (int k = 0; k < NC; k++) {
        xa[V_V + k] = fmax(xa[V_V + k], 0.0);
        st(i, O_XI + V_V + k, xa[V_V + k]);
      }
      on = lanes && do_diff;
      double z[NS], v[NC], dyv[NC];
#pragma unroll
      for (int k = 0; k < NS; k++) {
        z[k] = xa[V_Z + k] + (-1.0) * xb[V_Z + k];
        s[0] = fma(z[k], z[k], s[0]);
      }
#pragma unroll
      for (int k = 0; k < NX; k++) s[1] = fma(lc[k], lc[k], s[1]);
      if (i < N) {
#pragma unroll
        for (int k = 0; k < NX; k++) ln[k] = la[k] + (-1.0) * lb[k];
      }
#pragma unroll
      for (int k = 0; k < NC; k++) {
        v[k] = xa[V_V + k] + (-1.0) * xb[V_V + k];
        s[2] = fma(v[k], v[k], s[2]);
        const double yv = xa[V_Y + k] + (-1.0) * xb[V_Y + k];
        dyv[k] = yv + (-dat(i, D_d)[k * DS]);
      }
#pragma unroll
      for (int k = 0; k < NS; k++) st(i, O_DX + V_Z + k, z[k]);
#pragma unroll
      for (int k = 0; k < NX; k++) st(i, O_DX + V_L + k, lc[k]);
#pragma unroll
      for (int k = 0; k < NC; k++) {
        st(i, O_DX + V_V + k, v[k]);
        st(i, O_DX + V_Y + k, dyv[k]);
      }
      if (check) {
        const double* Qm = dat(i, D_Q);
        const double* Rm = dat(i, D_R);
        const double* Sm = dat(i, D_S);
        const double* Em = dat(i, D_E);
        const double* Lm = dat(i, D_L);
#pragma unroll
        for (int k = 0; k < NC; k++) {
          mx0 = fmax(mx0, Az_entry(z, i, k));
          mp1 = fmax(mp1, fabs(v[k]));
          sm1 += (-dat(i, D_d)[k * DS]) * v[k];
        }
#pragma unroll
        for (int rr = 0; rr < NX; rr++) {
          double gz, hh;
          if (i == 0) {
            gz = -z[rr];
            hh = -__ldg(x0 + rr);
          } else {
            const double* Ap = dat(i - 1, D_A);
            const double* Bp = dat(i - 1, D_B);
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) s1 = fma(Ap[(rr + cc * NX) * DS], zp[cc], s1);
#pragma unroll
            for (int cc = 0; cc < NU; cc++) s2 = fma(Bp[(rr + cc * NX) * DS], zp[NX + cc], s2);
            gz = (s1 + s2) - z[rr];
            hh = -dat(i - 1, D_c)[rr * DS];
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
            for (int cc = 0; cc < NX; cc++) s1 = fma(Qm[(rr + cc * NX) * DS], z[cc], s1);
#pragma unroll
            for (int cc = 0; cc < NU; cc++) s2 = fma(Sm[(cc + rr * NU) * DS], z[NX + cc], s2);
            fe = dat(i, D_q)[rr * DS];
            double sv = 0.0;
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(Em[(k + rr * NC) * DS], v[k], sv);
            p = sv + (-lc[rr]);
            if (i < N) {
              const double* Am = dat(i, D_A);
              double sa = 0.0;
#pragma unroll
              for (int cc = 0; cc < NX; cc++) sa = fma(Am[(cc + rr * NX) * DS], ln[cc], sa);
              p += sa;
            }
          } else {
            const int ru = rr - NX;
#pragma unroll
            for (int cc = 0; cc < NX; cc++) s1 = fma(Sm[(ru + cc * NU) * DS], z[cc], s1);
#pragma unroll
            for (int cc = 0; cc < NU; cc++) s2 = fma(Rm[(ru + cc * NU) * DS], z[NX + cc], s2);
            fe = dat(i, D_r)[ru * DS];
            double sv = 0.0;
#pragma unroll
            for (int k = 0; k < NC; k++) sv = fma(Lm[(k + ru * NC) * DS], v[k], sv);
            p = sv;
            if (i < N) {
              const double* Bm = dat(i, D_B);
              double sa = 0.0;
#pragma unroll
              for (int cc = 0; cc < NX; cc++) sa = fma(Bm[(cc + ru * NX) * DS], ln[cc], sa);
              p += sa;
            }
          }
          mx2 = fmax(mx2, fabs(s1 + s2));
          mx3 = fmax(mx3, fabs(z[rr]));
          sm0 += fe * z[rr];
          mp0 = fmax(mp0, fabs(p));
        }
      }
      // xk <- xi
#pragma unroll
      for (int k = 0; k < VSZ; k++) st(i, O_XK + k, xa[k]);
#pragma unroll
      for (int k = 0; k < NS; k++) zp[k] = z[k];
#pragma unroll
      for (int k = 0; k < NX; k++) lc[k] = ln[k];
    }
    on = lanes;
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

  // result block `from` (O_XK / O_XI / O_DX) -> the caller's instance-major arrays
  __device__ void write_result(int from, double* z, double* l, double* v, double* y) {
    for (int i = 0; i <= N; i++) {
      double x[VSZ];
#pragma unroll
      for (int k = 0; k < VSZ; k++) x[k] = ld(i, from + k);
#pragma unroll
      for (int k = 0; k < NS; k++) z[i * NS + k] = x[V_Z + k];
#pragma unroll
      for (int k = 0; k < NX; k++) l[i * NX + k] = x[V_L + k];
#pragma unroll
      for (int k = 0; k < NC; k++) {
        v[i * NC + k] = x[V_V + k];
        y[i * NC + k] = x[V_Y + k];
      }
    }
  }
};

// Does every instance carry the stage data of instance 0?  One pass over the
// inputs (bitwise comparison; x0 is per instance by definition).
__global__ void mpc_shared_detect(MpcData d, int batch, size_t sQ, size_t sR, size_t sS,
                                  size_t sq, size_t sr, size_t sA, size_t sB, size_t sc,
                                  size_t sE, size_t sL, size_t sd, int* mismatch) {
  const double* arr[11] = {d.Q, d.R, d.S, d.q, d.r, d.A, d.B, d.c, d.E, d.L, d.d};
  const size_t sz[11] = {sQ, sR, sS, sq, sr, sA, sB, sc, sE, sL, sd};
  bool bad = false;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  constexpr int CH = 64;  // instances are dealt to CH groups; a thread owns (offset, group)
  for (int k = 0; k < 11; k++) {
    const long long* p = (const long long*)arr[k];
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < sz[k] * CH; t += stride) {
      const size_t off = t % sz[k];
      const int group = (int)(t / sz[k]);
      const long long ref = p[off];
      for (int inst = 1 + group; inst < batch; inst += CH)
        if (p[(size_t)inst * sz[k] + off] != ref) bad = true;
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) *mismatch = 1;
}

// The common stage data, stage-major, in the layout Lane<..., true>::dat() reads.
template <int NX, int NU, int NC>
__global__ void mpc_shared_build(MpcData d, int N, const int* mismatch, double* sdata) {
  using LN = Lane<NX, NU, NC, true>;
  if (*mismatch) return;
  const double* src[11] = {d.Q, d.R, d.S, d.q, d.r, d.A, d.B, d.c, d.E, d.L, d.d};
  const int off[11] = {LN::D_Q, LN::D_R, LN::D_S, LN::D_q, LN::D_r, LN::D_A, LN::D_B, LN::D_c,
                       LN::D_E, LN::D_L, LN::D_d};
  const int cnt[11] = {NX * NX, NU * NU, NU * NX, NX, NU, NX * NX, NX * NU, NX,
                       NC * NX, NC * NU, NC};
  const int total = (N + 1) * LN::DSZ;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int i = e / LN::DSZ, o = e % LN::DSZ + LN::O_DAT;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const bool stage_only = (k == 5 || k == 6 || k == 7);  // A, B, c: N entries
      if (o >= off[k] && o < off[k] + cnt[k] && (!stage_only || i < N))
        v = src[k][(size_t)i * cnt[k] + (o - off[k])];
    }
    sdata[e] = v;
  }
}

template <int NX, int NU, int NC, bool SHARED>
__global__ void __launch_bounds__(32 * kLaneWarpsPerCta, 1) mpc_lane_kernel(const __grid_constant__ LaneArgs a) {
  using LN = Lane<NX, NU, NC, SHARED>;
  // exactly one of the two instantiations runs a launch pair
  if (a.mismatch != nullptr && ((*a.mismatch == 0) != SHARED)) return;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;  // warp in CTA
  const int warp = blockIdx.x * (blockDim.x >> 5) + wic;
  if (warp >= a.warps) return;  // (before any barrier: the whole warp leaves)
  LN p;
  p.N = a.N;
  p.nz = (a.N + 1) * LN::NS;
  p.nl = (a.N + 1) * NX;
  p.nv = (a.N + 1) * NC;
  p.ws = a.ws + (size_t)warp * a.ws_stride + lane;
  p.sdata = a.sdata;
  // shared memory: per warp [ slot mbarriers (128 bytes) | stage-block ring ], then ONE
  // copy of the common stage data for the CTA's warps (the warps are independent:
  // no CTA-wide barrier after this point)
  extern __shared__ __align__(128) unsigned char lane_smem[];
  constexpr size_t kWarpSmem = 128 + (size_t)LN::SLOTS * LN::RING_E * 256;
  unsigned char* my = lane_smem + (size_t)wic * kWarpSmem;
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < LN::SLOTS; m++) tma::mbar_init(tma::smem_addr(my) + 8 * m, 1);
    tma::fence_mbar_init();
  }
  p.bar0 = tma::smem_addr(my);
  p.ring = reinterpret_cast<double*>(my + 128) + lane;
  p.par = 0;
  p.on = true;
  p.bind(a, 0);  // lanes that never own an instance still execute the ring sweeps
  if (SHARED && a.sdata_smem) {
    double* sd = reinterpret_cast<double*>(lane_smem + (size_t)(blockDim.x >> 5) * kWarpSmem);
    const int total = (a.N + 1) * LN::DSZ;
    for (int e = threadIdx.x; e < total; e += blockDim.x) sd[e] = a.sdata[e];
    p.sdata = sd;
  }
  __syncthreads();

  lane_solve_loop(p, a);
}

typedef void (*LaneKernel)(const LaneArgs);
typedef void (*BuildKernel)(MpcData, int, const int*, double*);
struct LaneVariant {
  int nx, nu, nc;
  LaneKernel fn, fn_shared;
  BuildKernel build;
};
const LaneVariant kLaneVariants[] = {
    {4, 1, 4, mpc_lane_kernel<4, 1, 4, false>, mpc_lane_kernel<4, 1, 4, true>,
     mpc_shared_build<4, 1, 4>},
    {2, 1, 6, mpc_lane_kernel<2, 1, 6, false>, mpc_lane_kernel<2, 1, 6, true>,
     mpc_shared_build<2, 1, 6>},
};

}  // namespace

static int EnvLaneInt(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Batch size (in units of 16 instances per SM, the CTA kernel's default
// residency) from which the lane kernel is the faster one.  With every instance resident at once its time
// is the latency of one solve, whatever the batch; the CTA kernel needs one
// (several times shorter) wave per `capacity` instances.  Measured crossovers
// (tools/sweep_lane_min.py): servo 10 k instances = 4.2 x 2,368, double
// integrator 2.6 k = 1.1 x.
double MpcLaneCrossover(int nx, int nu, int nc) {
  if (nx == 4 && nu == 1 && nc == 4) return 4.2;
  if (nx == 2 && nu == 1 && nc == 6) return 1.1;
  return 4.0;
}

// dynamic shared memory of one warp: slot mbarriers (padded to 128 bytes) + the ring
size_t MpcLaneSmemBytes(int nx, int nu, int nc) {
  if (nx == 4 && nu == 1 && nc == 4)
    return 128 + (size_t)Lane<4, 1, 4>::SLOTS * Lane<4, 1, 4>::RING_E * 256;
  if (nx == 2 && nu == 1 && nc == 6)
    return 128 + (size_t)Lane<2, 1, 6>::SLOTS * Lane<2, 1, 6>::RING_E * 256;
  return 0;
}
// Warps per CTA (one CTA per SM): as many as the rings and one copy of the common
// stage data leave room for, at most kLaneWarpsPerCta.
static int LaneCtaWarps(int N, int nx, int nu, int nc, bool* sdata_smem) {
  const size_t ring = MpcLaneSmemBytes(nx, nu, nc);
  const size_t sd = MpcLaneSharedDoubles(N, nx, nu, nc) * 8;
  const size_t cap = 227 * 1024;
  int w = EnvLaneInt("FBSTAB_MPC_LANE_CTA_WARPS", kLaneWarpsPerCta);
  w = std::max(1, std::min(w, kLaneWarpsPerCta));
  bool in_smem = FBSTAB_LANE_SDATA_SMEM != 0;
  while (w > 1 && w * ring + (in_smem ? sd : 0) > cap) w--;
  if (w * ring + (in_smem ? sd : 0) > cap) in_smem = false;  // long horizons: read it through L1
  if (sdata_smem) *sdata_smem = in_smem;
  return w;
}
int MpcLaneWarpsPerSm(int N, int nx, int nu, int nc) {
  return LaneCtaWarps(N, nx, nu, nc, nullptr);
}
// Warps launched for `batch` instances on `sms` SMs (one CTA per SM): the batch is spread
// over every SM, and a lane's time is that of the instances it solves one after the
// other, so rather more warps than ceil(batch / 32) than idle SMs.
int MpcLaneWarps(int N, int nx, int nu, int nc, int batch, int sms, int* per_cta) {
  const int cta_warps = LaneCtaWarps(N, nx, nu, nc, nullptr);
  const int need = std::max(1, (batch + 31) / 32);
  const int ctas = std::min(sms, need);
  const int pc = std::min(cta_warps, (need + ctas - 1) / ctas);
  if (per_cta) *per_cta = pc;
  return ctas * pc;
}

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
  const size_t dat = 2 * (size_t)nx * nx + (size_t)nu * nu + 2 * (size_t)nu * nx + 2 * nx + nu +
                     (size_t)nc * nx + (size_t)nc * nu + nc;
  return 32 * (3 * (nz + nl + 2 * nv) + (nz + nl + nv) + 2 * nv + K * (fs + dat));
}

int MpcSharedDetect(int N, int nx, int nu, int nc, int batch, const MpcData& data,
                    int* mismatch, cudaStream_t stream) {
  const size_t K = N + 1;
  if (cudaMemsetAsync(mismatch, 0, sizeof(int), stream) != cudaSuccess) return 1;
  mpc_shared_detect<<<1184, 256, 0, stream>>>(
      data, batch, K * nx * nx, K * nu * nu, K * nu * nx, K * nx, K * nu, (size_t)N * nx * nx,
      (size_t)N * nx * nu, (size_t)N * nx, K * nc * nx, K * nc * nu, K * nc, mismatch);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

size_t MpcLaneSharedDoubles(int N, int nx, int nu, int nc) {
  const size_t dat = 2 * (size_t)nx * nx + (size_t)nu * nu + 2 * (size_t)nu * nx + 2 * nx + nu +
                     (size_t)nc * nx + (size_t)nc * nu + nc;
  return (size_t)(N + 1) * dat;
}

int MpcLaneLaunch(int N, int nx, int nu, int nc, int batch, int max_warps, const MpcData& data,
                  double* z, double* l, double* v, double* y, fbstab_out* out,
                  const fbstab_options& opts, double* ws, int* counter, int* mismatch,
                  double* sdata, cudaStream_t stream, bool shared_known) {
  const LaneVariant* var = nullptr;
  for (const LaneVariant& c : kLaneVariants)
    if (c.nx == nx && c.nu == nu && c.nc == nc) var = &c;
  if (!var) return 1;
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
  a.mismatch = nullptr;
  a.sdata = nullptr;
  bool sd_smem = false;
  const int cta_warps = LaneCtaWarps(N, nx, nu, nc, &sd_smem);
  const size_t smem_rings = (size_t)cta_warps * MpcLaneSmemBytes(nx, nu, nc);
  const size_t smem_shared = smem_rings + (sd_smem ? MpcLaneSharedDoubles(N, nx, nu, nc) * 8 : 0);
  a.sdata_smem = sd_smem ? 1 : 0;
  if (cudaFuncSetAttribute((const void*)var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem_rings) != cudaSuccess ||
      cudaFuncSetAttribute((const void*)var->fn_shared,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem_shared) != cudaSuccess)
    return 1;
  int dev = 0, sms = 148, per_cta = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int warps = std::min(max_warps, MpcLaneWarps(N, nx, nu, nc, batch, sms, &per_cta));
  const int ctas = (warps + per_cta - 1) / per_cta;
  a.warps = warps;
  if (mismatch && sdata) {
    // common-stage-data detection: one pass over the inputs, then exactly one of
    // the two kernels below does the work (no host round trip)
    // (shared_known: the caller passed ONE copy of the stage data -- no detection pass,
    // and only the shared-data kernel is launched)
    if (shared_known) {
      if (cudaMemsetAsync(mismatch, 0, sizeof(int), stream) != cudaSuccess) return 1;
    } else if (MpcSharedDetect(N, nx, nu, nc, batch, data, mismatch, stream)) {
      return 1;
    }
    var->build<<<32, 256, 0, stream>>>(data, N, mismatch, sdata);
    a.mismatch = mismatch;
    a.sdata = sdata;
    var->fn_shared<<<ctas, 32 * per_cta, smem_shared, stream>>>(a);
    if (shared_known) return cudaGetLastError() == cudaSuccess ? 0 : 1;
  } else if (shared_known) {
    return 1;  // the caller checks lane_sdata before asking for this path
  }
  var->fn<<<ctas, 32 * per_cta, smem_rings, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

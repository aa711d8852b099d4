// common.cuh -- device-side building blocks shared by every kernel:
// the cooperative "team" (one CTA per QP instance), deterministic reductions,
// the penalised Fischer-Burmeister function and its generalised gradient.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fbstab_b200.h"

namespace fbs {

constexpr int kMaxWarps = 32;
constexpr int kRedSlots = 8;  // values reduced at once

// Primal-dual iterate (z,l,v,y): reference fbstab/components/full_variable.h:31
struct Vars {
  double* z;
  double* l;
  double* v;
  double* y;
};

// One CTA cooperates on one QP instance.  blockDim.x is a multiple of 32.
struct Team {
  double* red;  // shared scratch, kMaxWarps*kRedSlots doubles
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
  __device__ __forceinline__ int nwarps() const { return blockDim.x >> 5; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// Optional per-phase cycle accounting (development builds only:
// -DFBSTAB_PHASE_TIMERS, tools/phase_timers.py).  lap(i) charges the cycles
// since the previous lap of this CTA to phase i.
#ifdef FBSTAB_PHASE_TIMERS
// (static: one copy per translation unit; only api.cu reads it back)
static __device__ unsigned long long g_phase_cycles[32];
static __device__ long long g_phase_last[1024];
__device__ __forceinline__ void phase_lap(int idx) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long now = clock64();
    atomicAdd(&g_phase_cycles[idx], (unsigned long long)(now - g_phase_last[blockIdx.x]));
    g_phase_last[blockIdx.x] = now;
  }
}
#define FBS_LAP(idx) ::fbs::phase_lap(idx)
#else
#define FBS_LAP(idx) ((void)0)
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;  // xor butterfly: every lane holds the bit-identical sum
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int off = 16; off; off >>= 1)
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// Sum N per-thread partials over the team; every thread receives identical
// results (fixed order -> deterministic, batch-size independent).
template <int N>
__device__ __forceinline__ void team_sum(const Team& t, double (&v)[N]) {
  static_assert(N <= kRedSlots, "too many values");
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = warp_sum(v[k]);
  const int nw = t.nwarps();
  if (nw == 1) return;
  __syncthreads();
  if (t.lane() == 0) {
#pragma unroll
    for (int k = 0; k < N; k++) t.red[t.warp() * kRedSlots + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += t.red[w * kRedSlots + k];
    v[k] = s;
  }
}
template <int N>
__device__ __forceinline__ void team_max(const Team& t, double (&v)[N]) {
  static_assert(N <= kRedSlots, "too many values");
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = warp_max(v[k]);
  const int nw = t.nwarps();
  if (nw == 1) return;
  __syncthreads();
  if (t.lane() == 0) {
#pragma unroll
    for (int k = 0; k < N; k++) t.red[t.warp() * kRedSlots + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = t.red[k];
    for (int w = 1; w < nw; w++) s = fmax(s, t.red[w * kRedSlots + k]);
    v[k] = s;
  }
}

// Penalised Fischer-Burmeister function, reference
// fbstab/components/full_residual.cc:115-118.
__device__ __forceinline__ double pfb(double a, double b, double alpha) {
  const double fb = a + b - sqrt(a * a + b * b);
  return alpha * fb + (1.0 - alpha) * fmax(0.0, a) * fmax(0.0, b);
}

// Penalised natural residual entry, full_residual.cc:92,103-104.
__device__ __forceinline__ double pnr(double y, double v, double alpha) {
  return alpha * fmin(y, v) + (1.0 - alpha) * fmax(0.0, y) * fmax(0.0, v);
}

// a / b from rb = RN(1 / b): one product and Markstein's exact-remainder
// correction give the correctly rounded quotient -- what `a / b` returns -- but
// quotients can share a reciprocal and, unlike the compiler's inline division, a
// zero or tiny quotient (v = 0 on every inactive constraint) does not leave the
// fast path for the out-of-line IEEE routine.  Where every thread or lane handles
// a different constraint that call was taken on practically every division
// (profiles/r1_dense_small_ncu_full_8warps.txt: eight calls per Newton step;
// the lane kernel: 10% of all stall samples on division lines).
__device__ __forceinline__ double div_r(double a, double b, double rb) {
  const double q = a * rb;
  const double rem = fma(-b, q, a);
  return fma(rem, rb, q);
}
__device__ __forceinline__ double div_nr(double a, double b) { return div_r(a, b, 1.0 / b); }

// Generalised gradient of the PFB function -> (gamma, mu):
// dense_cholesky_solver.cc:54-60,129-148 == riccati_linear_solver.cc:91-99,346-365.
__device__ __forceinline__ void pfb_barrier(double ys, double v, double alpha,
                                            double sigma, double* gamma,
                                            double* mu) {
  const double r = sqrt(ys * ys + v * v);
  double ga, gb;
  if (r < 1e-13) {  // zero_tolerance_
    const double d = 0.70710678118654752440;  // 1/sqrt(2)
    ga = alpha * (1.0 - d);
    gb = ga;
  } else {
    const double rr = 1.0 / r;
    const double qa = div_r(ys, r, rr), qb = div_r(v, r, rr);  // ys / r, v / r
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

}  // namespace fbs

// sparse_team.cu -- warp-per-instance FBstab for sparse QPs with a common pattern
// (FBstabSparse, second device path; sparse_lane.cu is the first).
//
// Why a second sparse kernel.  With one LANE per instance (sparse_lane.cu) an instance
// walks its 60 k value accesses per Newton round one after the other, each an
// indirectly addressed load from a 1.6 GB workspace, and the warp waits for the rounds
// of its slowest lane: 15 k solves/s on the servo OCP as a sparse QP, issue slots 4 %
// busy (profiles/r2_sparse_lane_ncu.txt).  Only 41 % of that chain is sequential by
// nature -- the rows of the up-looking LDL' and the triangular solves -- the rest
// (gather mat-vecs, barrier terms, assembly, element-wise sweeps) is independent per
// row or entry.  Here a WARP owns an instance (the engine.cuh state machine, the
// Problem policy below): the independent work is spread over the 32 lanes, and what is
// sequential runs out of SHARED memory -- the factor L, the reciprocal pivots and the
// work vector of the LDL' and of the triangular solves (nnz(L) + n doubles + nnz(L)
// 16-bit row indices per instance; servo OCP: 26.7 KB, eight instances per SM) -- at ~30 cycles per dependent step instead of an L2 / DRAM round
// trip.  Persistent single-warp CTAs pull instances from the global counter.
//
// The arithmetic is the lane kernel's operation for operation (same row sums in the
// same order, same schedule of the LDL', same forward / backward substitutions); only
// the norms are reduced across lanes (team_sum) instead of along one lane.
// Patterns whose factor does not fit shared memory take the lane kernel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.cuh"
#include "engine_args.cuh"
#include "sparse_lane.h"

namespace fbs {
namespace {

struct SparseTeamArgs {
  SparseDev d;
  const double *Hx, *f, *Gx, *h, *Ax, *b;
  CommonArgs c;
};

// Problem policy of engine.cuh over compressed-column data (reference concept:
// fbstab/components/abstract_components.h:24-62 + the LDL' wrapper
// tools/qdldl/qdldl_wrapper.h:19-84).
struct SparseProblem {
  int nz, nl, nv, n;
  SparseDev d;
  const double *Hx, *f, *Gx, *h, *Ax, *bvec;  // this instance
  // shared memory
  double *L, *yw, *xw;
  const unsigned short* Lis;  // the row indices of L (a 16-bit copy of d.Li: n < 65,536)
  // global workspace (per CTA); Dinv: the reciprocal pivots -- off the LDL' chain: a
  // visit's reciprocal is loaded with its schedule entry
  double *gamma, *mus, *sq, *r3, *tz, *Dinv, *Kx;

  __device__ __forceinline__ double b(int i) const { return bvec[i]; }
  __device__ __forceinline__ double fvec(int i) const { return f[i]; }
  __device__ __forceinline__ double hvec(int i) const { return h[i]; }

  __device__ double forcing_norm(const Team& t) const {  // dense_data.h:72-73
    double s[1] = {0.0};
    for (int i = t.rank(); i < nv; i += t.size()) s[0] += bvec[i] * bvec[i];
    for (int i = t.rank(); i < nz; i += t.size()) s[0] += f[i] * f[i];
    for (int i = t.rank(); i < nl; i += t.size()) s[0] += h[i] * h[i];
    team_sum(t, s);
    return sqrt(s[0]);
  }

  __device__ __forceinline__ double rowA(int k, const double* z) const {
    double s = 0.0;
    for (int q = d.Ar_ptr[k]; q < d.Ar_ptr[k + 1]; q++) s = fma(Ax[d.Ar_val[q]], z[d.Ar_col[q]], s);
    return s;
  }
  __device__ __forceinline__ double rowG(int r, const double* z) const {
    double s = 0.0;
    for (int q = d.Gr_ptr[r]; q < d.Gr_ptr[r + 1]; q++) s = fma(Gx[d.Gr_val[q]], z[d.Gr_col[q]], s);
    return s;
  }
  __device__ __forceinline__ double rowH(int i, const double* z) const {
    double s = 0.0;
    for (int q = d.Hr_ptr[i]; q < d.Hr_ptr[i + 1]; q++) s = fma(Hx[d.Hr_val[q]], z[d.Hr_col[q]], s);
    return s;
  }
  __device__ __forceinline__ double colA(int i, const double* v) const {  // (A'v)_i
    double s = 0.0;
    for (int e = d.Ap[i]; e < d.Ap[i + 1]; e++) s = fma(Ax[e], v[d.Ai[e]], s);
    return s;
  }
  __device__ __forceinline__ double colG(int i, const double* l) const {  // (G'l)_i
    double s = 0.0;
    for (int e = d.Gp[i]; e < d.Gp[i + 1]; e++) s = fma(Gx[e], l[d.Gi[e]], s);
    return s;
  }

  // y = b - A z (full_variable.cc:47-53)
  __device__ void margin(const Team& t, const double* z, double* y) const {
    for (int k = t.rank(); k < nv; k += t.size()) y[k] = bvec[k] - rowA(k, z);
    t.sync();
  }

  // tz = ((f + Hz) + G'l) + A'v ; tl = h - Gz (full_residual.cc:52-63)
  __device__ void kkt(const Team& t, const Vars& x, double* oz, double* ol) const {
    for (int i = t.rank(); i < nz; i += t.size())
      oz[i] = ((f[i] + rowH(i, x.z)) + colG(i, x.l)) + colA(i, x.v);
    for (int r = t.rank(); r < nl; r += t.size()) ol[r] = h[r] - rowG(r, x.z);
    t.sync();
  }

  __device__ __forceinline__ double kval(int e, double sigma) const {
    const int kind = d.Kkind[e], idx = d.Kidx[e];
    if (kind == KSRC_H) return Hx[idx];
    if (kind == KSRC_H_SIGMA) return Hx[idx] + sigma;
    if (kind == KSRC_SIGMA) return sigma;
    if (kind == KSRC_G) return Gx[idx];
    if (kind == KSRC_NEG_SIGMA) return -sigma;
    if (kind == KSRC_A) return sq[d.Krow[e]] * Ax[idx];
    return -1.0;
  }

  // LinearSolver::Initialize: barrier terms, then the up-looking LDL' of the permuted
  // K (the schedule of QDLDL_factor), row by row.  False on a zero / NaN pivot.
  __device__ bool factor(const Team& t, const Vars& x, const Vars& xbar, double sigma,
                         double alpha) {
    for (int k = t.rank(); k < nv; k += t.size()) {
      const double ys = x.y[k] + sigma * (x.v[k] - xbar.v[k]);
      double ga, mu;
      pfb_barrier(ys, x.v[k], alpha, sigma, &ga, &mu);
      gamma[k] = ga;
      mus[k] = mu;
      sq[k] = sqrt(div_nr(ga, mu));
    }
    for (int i = t.rank(); i < n; i += t.size()) yw[i] = 0.0;
    t.sync();
    bool ok = true;
    const int lane = t.rank();
    // The rows are a chain (row k needs the pivots and columns before it), so what is on
    // the chain is kept short: the entries of a row's column of K are fetched two rows
    // ahead (lane p holds entry p); a row's schedule (column, slot, start of the column)
    // is loaded by the lanes side by side and handed round by shuffles; L, its row
    // indices, the reciprocal pivots and y are in shared memory; the warp synchronises
    // with __syncwarp (the CTA is one warp).
    // K in one parallel pass (independent gathers, 32 in flight) ...
    for (int e = lane; e < d.nnzK; e += 32) Kx[e] = kval(e, sigma);
    __syncwarp();
    // ... and the rows read it two rows ahead
    auto fetch = [&](int k, double* diag) {
      double v = 0.0;
      if (k < n) {
        const int p0 = d.Kp[k], p1 = d.Kp[k + 1];
        *diag = Kx[p1 - 1];
        if (p0 + lane < p1 - 1) v = Kx[p0 + lane];
      }
      return v;
    };
    double dA = 0.0, dB = 0.0;
    double kA = fetch(0, &dA), kB = fetch(1, &dB);
    for (int k = 0; k < n; k++) {
      // scatter column k of K (its diagonal entry is the last one: rows are sorted)
      const int p0 = d.Kp[k], p1 = d.Kp[k + 1];
      if (p0 + lane < p1 - 1) yw[d.Ki[p0 + lane]] = kA;
      for (int p = p0 + 32 + lane; p < p1 - 1; p += 32) yw[d.Ki[p]] = Kx[p];
      double dk = dA;
      kA = kB;
      dA = dB;
      kB = fetch(k + 2, &dB);
      __syncwarp();
      const int q0 = d.Sp[k], q1 = d.Sp[k + 1];
      for (int qb = q0; qb < q1; qb += 32) {
        // this lane's entry of the schedule
        int mc = 0, mslot = 0, mj0 = 0;
        double mdi = 0.0;
        if (qb + lane < q1) {
          mc = d.Sc[qb + lane];
          mslot = d.St[qb + lane];
          mj0 = d.Lp[mc];
          mdi = Dinv[mc];  // (columns visited by row k are < k: written in earlier rows)
        }
        const int cnt = min(32, q1 - qb);
        for (int u = 0; u < cnt; u++) {
          const int c = __shfl_sync(0xffffffffu, mc, u);
          const int slot = __shfl_sync(0xffffffffu, mslot, u);
          const int j0 = __shfl_sync(0xffffffffu, mj0, u);
          const double di = __shfl_sync(0xffffffffu, mdi, u);
          const double yc = yw[c];
          __syncwarp();  // every lane holds yc before y is updated
          for (int j = j0 + lane; j < slot; j += 32) {
            const int r = Lis[j];
            yw[r] = fma(-L[j], yc, yw[r]);
          }
          const double lx = yc * di;
          dk = fma(-yc, lx, dk);
          if (lane == 0) {
            L[slot] = lx;
            yw[c] = 0.0;
          }
          __syncwarp();
        }
      }
      if (!(fabs(dk) > 0.0)) ok = false;
      if (lane == 0) Dinv[k] = 1.0 / dk;
      __syncwarp();
    }
    return ok;
  }

  // LinearSolver::Solve on r = -(rz, rl, rv) -> dx (QDLDL_solve: L, D^-1, L')
  __device__ void solve(const Team& t, const double* rz, const double* rl, const double* rv,
                        const Vars& dx) {
    const int lane = t.rank();
    for (int k = lane; k < nv; k += t.size()) r3[k] = div_nr(-rv[k], mus[k]);
    t.sync();
    for (int i = lane; i < nz; i += t.size()) xw[d.iperm[i]] = (-rz[i]) - colA(i, r3);
    for (int r = lane; r < nl; r += t.size()) xw[d.iperm[nz + r]] = rl[r];
    for (int k = lane; k < nv; k += t.size()) xw[d.iperm[nz + nl + k]] = 0.0;
    t.sync();
    for (int i = 0; i < n; i++) {
      const int j0 = d.Lp[i], j1 = d.Lp[i + 1];
      if (j0 == j1) continue;
      const double xi = xw[i];
      __syncwarp();
      for (int j = j0 + lane; j < j1; j += 32) {
        const int r = Lis[j];
        xw[r] = fma(-L[j], xi, xw[r]);
      }
      __syncwarp();
    }
    for (int i = lane; i < n; i += 32) xw[i] = xw[i] * Dinv[i];
    __syncwarp();
    // backward substitution: the sums run along one lane in the order of QDLDL_solve
    if (lane == 0) {
      int j1 = d.Lp[n];
      for (int i = n - 1; i >= 0; i--) {
        const int j0 = d.Lp[i];
        double xi = xw[i];
        for (int j = j0; j < j1; j++) xi = fma(-L[j], xw[Lis[j]], xi);
        xw[i] = xi;
        j1 = j0;
      }
    }
    t.sync();
    for (int i = lane; i < nz; i += t.size()) dx.z[i] = xw[d.iperm[i]];
    for (int r = lane; r < nl; r += t.size()) dx.l[r] = xw[d.iperm[nz + r]];
    t.sync();
    // dv = (rv + gamma .* (A dz)) ./ mu ; dy = b - A dz
    for (int k = lane; k < nv; k += t.size()) {
      const double s = rowA(k, dx.z);
      dx.v[k] = div_nr(gamma[k] * s + (-rv[k]), mus[k]);
      dx.y[k] = bvec[k] - s;
    }
    t.sync();
  }

  // FullFeasibility::CheckFeasibility (full_feasibility.cc:25-88)
  __device__ int feasibility(const Team& t, const Vars& dx, double tol) {
    double mx[4] = {-INFINITY, 0.0, 0.0, 0.0};
    double sm[2] = {0.0, 0.0};
    double mp[3] = {0.0, 0.0, 0.0};
    for (int k = t.rank(); k < nv; k += t.size()) {
      mx[0] = fmax(mx[0], rowA(k, dx.z));
      mp[1] = fmax(mp[1], fabs(dx.v[k]));
      sm[1] += bvec[k] * dx.v[k];
    }
    for (int r = t.rank(); r < nl; r += t.size()) {
      mx[1] = fmax(mx[1], fabs(rowG(r, dx.z)));
      mp[2] = fmax(mp[2], fabs(dx.l[r]));
      sm[1] += h[r] * dx.l[r];
    }
    for (int i = t.rank(); i < nz; i += t.size()) {
      mx[2] = fmax(mx[2], fabs(rowH(i, dx.z)));
      mx[3] = fmax(mx[3], fabs(dx.z[i]));
      sm[0] += f[i] * dx.z[i];
      mp[0] = fmax(mp[0], fabs(colA(i, dx.v) + colG(i, dx.l)));
    }
    team_max(t, mx);
    team_max(t, mp);
    team_sum(t, sm);
    t.sync();
    const double w = mx[3];
    const bool dual_infeasible = (mx[0] <= w * tol) && (mx[1] <= tol * w) &&
                                 (mx[2] <= tol * w) && (sm[0] < 0.0) && (w > 1e-14);
    const double u = fmax(mp[1], mp[2]);
    const bool primal_infeasible = (mp[0] <= tol * u) && (sm[1] < 0.0);
    return (primal_infeasible ? 1 : 0) + (dual_infeasible ? 2 : 0);
  }
};

__device__ inline double* Carve(double*& p, size_t n) {
  double* r = p;
  p += n;
  return r;
}

__global__ void __launch_bounds__(32, 8) sparse_team_kernel(const __grid_constant__ SparseTeamArgs a) {
  extern __shared__ double dyn_smem[];
  __shared__ double red[kRedSlots];  // one warp: team_sum / team_max use slot row 0 only
  __shared__ int s_inst;
  Team t{red};
  const CommonArgs& c = a.c;
  const SparseDev& d = a.d;
  const int nz = d.nz, nl = d.nl, nv = d.nv, n = d.n;
  double* ws0 = c.ws + (size_t)blockIdx.x * c.ws_stride;
  {  // the pattern of L next to its values: one copy per CTA for all its instances
    unsigned short* lis = reinterpret_cast<unsigned short*>(dyn_smem + d.nnzL + (size_t)n);
    for (int j = threadIdx.x; j < d.nnzL; j += blockDim.x) lis[j] = (unsigned short)d.Li[j];
    __syncthreads();
  }
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(c.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= c.batch) break;
    double* ws = ws0;
    Buffers w;
    Vars* vs[4] = {&w.xk, &w.xi, &w.xp, &w.dx};
    for (int k = 0; k < 4; k++) {
      vs[k]->z = Carve(ws, nz);
      vs[k]->l = Carve(ws, nl);
      vs[k]->v = Carve(ws, nv);
      vs[k]->y = Carve(ws, nv);
    }
    w.ri.z = Carve(ws, nz);
    w.ri.l = Carve(ws, nl);
    w.ri.v = Carve(ws, nv);
    SparseProblem p;
    p.nz = nz;
    p.nl = nl;
    p.nv = nv;
    p.n = n;
    p.d = d;
    p.Hx = a.Hx + (size_t)inst * d.nnzH;
    p.f = a.f + (size_t)inst * nz;
    p.Gx = a.Gx + (size_t)inst * d.nnzG;
    p.h = a.h + (size_t)inst * nl;
    p.Ax = a.Ax + (size_t)inst * d.nnzA;
    p.bvec = a.b + (size_t)inst * nv;
    p.gamma = Carve(ws, nv);
    p.mus = Carve(ws, nv);
    p.sq = Carve(ws, nv);
    p.r3 = Carve(ws, nv);
    p.tz = Carve(ws, nz);
    p.Dinv = Carve(ws, n);
    p.Kx = Carve(ws, d.nnzK);
    double* sm = dyn_smem;
    p.L = Carve(sm, d.nnzL);
    p.yw = Carve(sm, n);
    p.xw = p.yw;  // the LDL' work vector is all zeros again when factor() returns
    p.Lis = reinterpret_cast<const unsigned short*>(sm);  // filled once per CTA, above
    solve_instance(t, p, c.opts, w, c.z + (size_t)inst * nz, c.l + (size_t)inst * nl,
                   c.v + (size_t)inst * nv, c.y + (size_t)inst * nv, c.out + inst);
  }
}

}  // namespace

size_t SparseTeamSmemBytes(const SparseDev& d) {
  return sizeof(double) * ((size_t)d.nnzL + (size_t)d.n) +
         sizeof(unsigned short) * (((size_t)d.nnzL + 3) & ~(size_t)3);
}
size_t SparseTeamWsDoubles(const SparseDev& d) {
  const size_t vs = (size_t)d.nz + d.nl + 2 * (size_t)d.nv;
  return 4 * vs + ((size_t)d.nz + d.nl + d.nv) + 4 * (size_t)d.nv + d.nz + d.n + d.nnzK;
}

// Resident single-warp CTAs per SM, 0 when the factor does not fit shared memory
// (or FBSTAB_SPARSE_TEAM=0): the caller then takes the lane kernel.
int SparseTeamCtasPerSm(const SparseDev& d) {
  if (const char* e = getenv("FBSTAB_SPARSE_TEAM"))
    if (atoi(e) == 0) return 0;
  const size_t smem = SparseTeamSmemBytes(d);
  if (smem > 100 * 1024 || d.n > 65535) return 0;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute((const void*)sparse_team_kernel,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)sparse_team_kernel, 32,
                                                    smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return occ;
}

int SparseTeamLaunch(const SparseDev& d, int batch, int ctas, const double* Hx, const double* f,
                     const double* Gx, const double* h, const double* Ax, const double* b,
                     double* z, double* l, double* v, double* y, fbstab_out* out,
                     const fbstab_options& opts, double* ws, int* counter, cudaStream_t stream) {
  SparseTeamArgs a;
  a.d = d;
  a.Hx = Hx;
  a.f = f;
  a.Gx = Gx;
  a.h = h;
  a.Ax = Ax;
  a.b = b;
  a.c.batch = batch;
  a.c.z = z;
  a.c.l = l;
  a.c.v = v;
  a.c.y = y;
  a.c.out = out;
  a.c.ws = ws;
  a.c.ws_stride = SparseTeamWsDoubles(d);
  a.c.counter = counter;
  a.c.vec_in_smem = 0;
  a.c.opts = opts;
  a.c.comp = -1;
  memset(&a.c.io, 0, sizeof(a.c.io));
  sparse_team_kernel<<<std::min(ctas, batch), 32, SparseTeamSmemBytes(d), stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

// dense_large.cuh -- Problem policy for LARGE dense QPs (BASELINE config 5:
// nz=512, nl=128, nv=1024): one 256-thread CTA per instance, two CTAs per SM
// (one instance's latency-bound phases overlap the other's DMMA phases), the KKT reduction
// and its factorisation on the FP64 tensor cores.
//
// Follows DenseCholeskySolver (reference dense_cholesky_solver.cc:32-127).
// The reference factors K = [E G'; G -sigma I], E = H + sigma I + A' Gamma A,
// with Eigen's diagonally pivoted LDL'; that pivot rule eliminates the E block
// first (see dense_problem.cuh), so the same elimination is done here as
//     E = L L'            blocked right-looking Cholesky (NB = 64)
//     W' = G L^-T         falls out of the panel solves (the G rows are simply
//                         further rows of every panel)
//     S = sigma I + W'W   accumulated by the same trailing updates with the
//                         sign flipped, then S = Ls Ls'
// which is the oracle's `variant 2`.  What is a genuine dense contraction runs
// as FP64 DMMA (mma.sync.m8n8k4.f64):
//   * A' Gamma A : 128x128 tiles of E, K = nv deep (75% of the flops);
//   * every trailing update of the blocked Cholesky (K = 64 deep).
// Operand chunks (128 x 16) are staged global -> shared by cp.async through a
// three-deep ring, two chunks ahead of the DMMAs.  Diagonal
// blocks and panel solves work in shared memory; the triangular solves of
// ::Solve are blocked the same way (one warp solves a diagonal block with
// shuffles, all warps apply the block column).
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is looked up at run time)

#include "dense_problem.cuh"
#include "tma.cuh"

namespace fbs {
namespace dl {

constexpr int kThreads = 256;  // two CTAs (= two instances) per SM
constexpr int TB = 128;   // tile rows of the DMMA products
constexpr int TBN = 64;   // tile columns
constexpr int KC = 16;   // depth of one staged chunk
constexpr int KP = 20;   // padded depth stride in shared memory (conflict-free fragments)
constexpr int NB = 64;   // Cholesky block
constexpr int DP = NB + 1;
constexpr int TR = 128;  // rows of one panel-solve tile
// shared-memory carve (doubles)
constexpr int kStages = 3;                      // cp.async ring depth
constexpr int kStageDoubles = (TB + TBN) * KP + KC;  // two operand chunks + Gamma
constexpr int kStage = kStages * kStageDoubles;
constexpr int kDiag = NB * DP + NB;             // diagonal block + its pivots
constexpr int kPanel = TR * NB;
constexpr int kSmemDoubles = (kDiag + kPanel) > kStage ? (kDiag + kPanel) : kStage;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// 8-byte asynchronous copy global -> shared (LDGSTS); `valid` false zero-fills.
__device__ __forceinline__ void cp8(double* dst, const double* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMA operand staging for A' Gamma A ---------------------------------------
// One 2-D tensor map covers A of the whole batch (dim0 = k, contiguous, nv long;
// dim1 = the batch * nz columns of A); a box is 16 k x 64 columns = 64 rows of
// 128 bytes, 128-byte swizzled, so that the DMMA fragment reads are the
// 2-wavefront minimum without padding.  One thread issues 2-3 boxes per chunk;
// the other 255 issue no copy instruction at all (every LDGSTS of the cp.async
// path costs the FP64 pipe ~35 cycles, tools/probes/dmma_probe.cu: 21.6 ->
// 17.9 cycles per DMMA with boxes).
constexpr int kTmaBoxBytes = 64 * 128;
constexpr int kTmaStageBytes = (TB + TBN) / 64 * kTmaBoxBytes;
static_assert(kStages * kTmaStageBytes <= kSmemDoubles * 8, "TMA ring must fit the scratch");
struct SmemHeader {  // first bytes of the dynamic shared memory
  unsigned long long full[kStages];  // mbarriers: chunk landed
  unsigned seq;                      // chunks issued so far (slot / parity bookkeeping)
};
__device__ __forceinline__ void tma_box(unsigned dst, const void* tmap, int c0, int c1,
                                        unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}

// Tile product with TMA-staged operands, both k-contiguous in global memory:
//   MODE 1: acc += sum_k A(k, rowI + i) Gam(k) A(k, rowJ + j)       (A' Gamma A)
//   MODE 2: acc += csgn(j) sum_k Xt(rowI + i, k) Xt(rowJ + j, k)    (trailing update; Xt is
//           the row-major copy of the current panel that the panel solve leaves behind)
// rowI / rowJ: dim-1 coordinates of the tile's first I / J operand row in the
// tensor map.  seq: running chunk number (uniform over the CTA).
template <int MODE>
__device__ __forceinline__ void tile_tma(double (&acc)[4][4][2], unsigned ring,
                                         unsigned bar0, const void* tmap, int rowI,
                                         int rowJ, int j_alias, bool active,
                                         const double* Gam, const double (&csgn)[4],
                                         int depth, unsigned& seq) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r8 = lane >> 2, c4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int nchunk = (depth + KC - 1) / KC;
  auto issue = [&](int ch, unsigned sq) {
    const unsigned slot = sq % kStages;
    const unsigned bar = bar0 + 8 * slot;
    const unsigned dst = ring + slot * kTmaStageBytes;
    tma::mbar_arrive_expect_tx(bar, (j_alias < 0 ? (TB + TBN) : TB) * 128);
#pragma unroll
    for (int bx = 0; bx < TB / 64; bx++)
      tma_box(dst + bx * kTmaBoxBytes, tmap, ch * KC, rowI + 64 * bx, bar);
    if (j_alias < 0) tma_box(dst + (TB / 64) * kTmaBoxBytes, tmap, ch * KC, rowJ, bar);
  };
  if (tid == 0) {
    issue(0, seq);
    if (nchunk > 1) issue(1, seq + 1);
  }
  for (int ch = 0; ch < nchunk; ch++) {
    const unsigned sq = seq + ch, slot = sq % kStages;
    double gk[KC / 4];
    if (MODE == 1) {
#pragma unroll
      for (int kk = 0; kk < KC / 4; kk++) {
        const int k = ch * KC + 4 * kk + c4;
        gk[kk] = (k < depth) ? __ldg(Gam + k) : 0.0;
      }
    }
    tma::mbar_wait(bar0 + 8 * slot, (sq / kStages) & 1);  // chunk ch has landed
    __syncthreads();                                       // everyone is done with chunk ch-1
    if (tid == 0 && ch + 2 < nchunk) issue(ch + 2, sq + 2);  // into the slot chunk ch-1 used
    if (!active) continue;
    const unsigned SI = ring + slot * kTmaStageBytes;
    const unsigned SJ = (j_alias >= 0) ? SI + j_alias * 128 : SI + (TB / 64) * kTmaBoxBytes;
#pragma unroll
    for (int kk = 0; kk < KC / 4; kk++) {
      // element (row, k) of a box sits at row*128 + (((k/2) ^ (row%8)) * 16) + (k%2)*8
      const unsigned koff = ((unsigned)((2 * kk + (c4 >> 1)) ^ r8) << 4) | ((unsigned)(c4 & 1) << 3);
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; a++) af[a] = lds64(SI + (32 * wm + 8 * a + r8) * 128 + koff);
#pragma unroll
      for (int b = 0; b < 4; b++)
        bf[b] = (MODE == 1 ? gk[kk] : csgn[b]) * lds64(SJ + (32 * wn + 8 * b + r8) * 128 + koff);
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  seq += nchunk;
  __syncthreads();  // the slots are free for the next tile
}

// acc(128x128, this warp's 32x32 part) += sum_k opI(i,k) * w * opJ(j,k), where
// the weight w is 1 (MODE 0), scale[k] (MODE 1: Gamma of A' Gamma A) or
// csgn[b] = +-1 per 8-column group of the tile (MODE 2: the trailing update
// subtracts inside E and W, adds inside S).  The caller PRELOADS acc with the
// tile's current value (H + sigma I, or the trailing K): those 32 global loads
// per thread are all in flight while the first operand chunks arrive, instead
// of 32 dependent load -> store round trips after the product.
// AddrI / AddrJ: (idx in [0,128), k in [0,depth)) -> global address of the
// element, or nullptr outside the matrix (zero filled).  A tile is 128 x 64;
// the 8 warps form a 4x2 grid, warp (wm,wn) owns rows 32wm.., cols 32wn.. as
// 4x4 m8n8 DMMA tiles.  `j_alias` >= 0: the J operand rows are rows
// j_alias.. of the I operand (tiles on the diagonal), nothing extra is staged.
// `active` false: this warp's 32x32 part lies strictly above the diagonal; it
// only helps staging.  (Measured and rejected, tools/probes/dmma_probe.cu and
// profiles/r1_dense_large_phases.txt: 16-byte cp.async with a row-pair thread
// mapping, a k-major trailing layout, TMA bulk row copies, register staging.)
// Operand chunks travel global -> shared
// as cp.async copies through a kStages-deep ring, two chunks ahead of the
// DMMAs, with one barrier per chunk.
template <int MODE, class AddrI, class AddrJ>
__device__ __forceinline__ void mma_tile(double (&acc)[4][4][2], double* sm, int depth,
                                         AddrI ai, AddrJ aj, const double* scale, int j_alias,
                                         bool active, const double (&csgn)[4]) {
  constexpr bool SCALED = (MODE == 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r8 = lane >> 2, c4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int sidx = 8 * warp + r8;  // staging: this thread's tile row/col index (0..63)
  const int nchunk = (depth + KC - 1) / KC;
  auto issue = [&](int ch) {
    if (ch < nchunk) {
      const int k0 = ch * KC;
      double* SI = sm + (ch % kStages) * kStageDoubles;
      double* SJ = SI + TB * KP;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int k = k0 + 4 * q + c4;
#pragma unroll
        for (int hh = 0; hh < TB / 64; hh++) {
          const int idx = 64 * hh + sidx;
          const double* pi = (k < depth) ? ai(idx, k) : nullptr;
          cp8(SI + idx * KP + 4 * q + c4, pi ? pi : scale, pi != nullptr);
        }
        if (j_alias < 0) {
          const double* pj = (k < depth) ? aj(sidx, k) : nullptr;
          cp8(SJ + sidx * KP + 4 * q + c4, pj ? pj : scale, pj != nullptr);
        }
      }
      if (SCALED && tid < KC)
        cp8(SI + (TB + TBN) * KP + tid, scale + (k0 + tid < depth ? k0 + tid : 0),
            k0 + tid < depth);
    }
    cp_commit();
  };
  issue(0);
  issue(1);
  for (int ch = 0; ch < nchunk; ch++) {
    cp_wait<1>();     // this thread's copies of chunk ch have landed
    __syncthreads();  // everyone's have; everyone is done with chunk ch-1
    issue(ch + 2);    // into the buffer chunk ch-1 used
    if (!active) continue;  // sub-tile strictly above the diagonal: nothing to compute
    const double* SI = sm + (ch % kStages) * kStageDoubles;
    const double* SJ = (j_alias >= 0) ? SI + j_alias * KP : SI + TB * KP;
    const double* SG = SI + (TB + TBN) * KP;
#pragma unroll
    for (int kk = 0; kk < KC / 4; kk++) {
      double af[4], bf[4];
      const double g = SCALED ? SG[4 * kk + c4] : 1.0;
#pragma unroll
      for (int a = 0; a < 4; a++) af[a] = SI[(32 * wm + 8 * a + r8) * KP + 4 * kk + c4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double v = SJ[(32 * wn + 8 * b + r8) * KP + 4 * kk + c4];
        bf[b] = SCALED ? g * v : (MODE == 2 ? csgn[b] * v : v);
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  cp_wait<0>();
  __syncthreads();  // the ring is free for the next tile (or the next phase)
}

// Visits this thread's accumulator entries: f(tile_row, tile_col, value).
template <class F>
__device__ __forceinline__ void for_each_acc(const double (&acc)[4][4][2], F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r8 = lane >> 2, c4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int r = 32 * wm + 8 * a + r8, c = 32 * wn + 8 * b + 2 * c4;
      f(r, c, acc[a][b][0]);
      f(r, c + 1, acc[a][b][1]);
    }
}

// s + sum_k a[k*sa] * x[k*sx], k = 0..n-1, accumulated in order (one chain, so
// bit-identical to the plain loop) with eight loads of `a` in flight: the
// streaming mat-vecs read H, G and A straight from HBM and are latency-bound.
__device__ __forceinline__ double dot_stream(const double* __restrict__ a, size_t sa,
                                             const double* x, int sx, int n, double s) {
  int k = 0;
  for (; k + 8 <= n; k += 8) {
    double t[8];
#pragma unroll
    for (int u = 0; u < 8; u++) t[u] = __ldg(a + (size_t)(k + u) * sa);
#pragma unroll
    for (int u = 0; u < 8; u++) s = fma(t[u], x[(k + u) * sx], s);
  }
  for (; k < n; k++) s = fma(__ldg(a + (size_t)k * sa), x[k * sx], s);
  return s;
}

// Preloads this thread's accumulator entries: acc = f(tile_row, tile_col).
template <class F>
__device__ __forceinline__ void init_acc(double (&acc)[4][4][2], F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r8 = lane >> 2, c4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int r = 32 * wm + 8 * a + r8, c = 32 * wn + 8 * b + 2 * c4;
      acc[a][b][0] = f(r, c);
      acc[a][b][1] = f(r, c + 1);
    }
}

}  // namespace dl

struct DenseLargeProblem : DenseProblem {
  double* rpiv = nullptr;  // n reciprocal pivots 1 / L(j,j) of the current factor (global)
  double* sm;  // dl::kSmemDoubles of dynamic shared memory (1024-byte aligned)
  dl::SmemHeader* hdr = nullptr;  // mbarriers + chunk counter of the TMA ring
  const void* tmA = nullptr;      // batch-wide tensor map of A, or nullptr (cp.async staging)
  int tma_row0 = 0;               // dim-1 coordinate of this instance's first column of A
  const void* tmX = nullptr;      // tensor map of the row-major panel copies, or nullptr
  double* xt = nullptr;           // this CTA's panel copy Xt[r * NB + k], r = row of K
  int xt_row0 = 0;                // dim-1 coordinate of this CTA's row 0 in tmX

  // DenseProblem::kkt / ::margin with eight loads in flight per thread (same
  // sums in the same order as the base class)
  __device__ void kkt(const Team& t, const Vars& x, double* oz, double* ol) const {
    FBS_LAP(15);
    for (int i = t.rank(); i < n; i += t.size()) {
      if (i < nz) {
        oz[i] = f[i] + dl::dot_stream(H + i, (size_t)nz, x.z, 1, nz, 0.0);
      } else {
        const int k = i - nz;
        ol[k] = h[k] - dl::dot_stream(G + k, (size_t)nl, x.z, 1, nz, 0.0);
      }
    }
    t.sync();
    FBS_LAP(11);
    const int lane = t.lane();
    for (int i = t.warp(); i < nz; i += t.nwarps()) {
      double s1 = (lane < nl) ? dl::dot_stream(G + (size_t)i * nl + lane, 32, x.l + lane, 32,
                                               (nl - lane + 31) / 32, 0.0)
                              : 0.0;
      double s2 = (lane < nv) ? dl::dot_stream(A + (size_t)i * nv + lane, 32, x.v + lane, 32,
                                               (nv - lane + 31) / 32, 0.0)
                              : 0.0;
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) oz[i] = (oz[i] + s1) + s2;
    }
    t.sync();
    FBS_LAP(12);
  }
  __device__ void margin(const Team& t, const double* z, double* y) const {
    for (int i = t.rank(); i < nv; i += t.size())
      y[i] = bvec[i] - dl::dot_stream(A + i, (size_t)nv, z, 1, nz, 0.0);
    t.sync();
  }

  // In-place lower Cholesky of the bs x bs diagonal block at (c0,c0) of K,
  // right-looking in shared memory; the factor is left in D (stride DP) and
  // written back.  false on a pivot <= 0 (Eigen LLT's failure rule).
  // bs == NB fast path: thread (i, q) = (tid % 64, tid / 64) keeps the entries
  // (i, q + 4m) of the block in registers; a step publishes the scaled column
  // through shared memory and every thread updates its own entries.  No
  // read-modify-write of shared memory, two light barriers per step, and the
  // same operations in the same order as the general path (bit-identical).
  __device__ __noinline__ bool factor_diag_full(int c0, double* D, double* dg) {
    constexpr int BS = dl::NB, NQ = dl::kThreads / 64, NM = BS / NQ;
    const int tid = threadIdx.x, i = tid & 63, q = tid >> 6;
    double a[NM];
#pragma unroll
    for (int m = 0; m < NM; m++) {
      const int k = q + NQ * m;
      a[m] = (k <= i) ? K[(c0 + i) + (size_t)(c0 + k) * n] : 0.0;
    }
    double* col = D;  // the block itself is only assembled in D after the loop
    bool ok = true;
#pragma unroll
    for (int j = 0; j < BS; j++) {
      constexpr int dummy = 0;
      (void)dummy;
      const int qj = j % NQ, mj = j / NQ;
      if (i == j && q == qj) {
        a[mj] = sqrt(a[mj]);
        dg[0] = a[mj];
      }
      __syncthreads();
      const double sd = dg[0];
      if (!(sd > 0.0)) ok = false;  // pivot <= 0 or NaN
      if (q == qj && i > j) {
        a[mj] = a[mj] / sd;
        col[i] = a[mj];
      }
      __syncthreads();
      if (i > j) {
        const double lij = col[i];
#pragma unroll
        for (int m = 0; m < NM; m++) {
          if (NQ * m + NQ - 1 > j) {  // some thread group still has column q + NQ*m > j
            const int k = q + NQ * m;
            if (k > j && k <= i) a[m] = fma(-lij, col[k], a[m]);
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NM; m++) {
      const int k = q + NQ * m;
      D[i + k * dl::DP] = (k <= i) ? a[m] : 0.0;
      if (k <= i) K[(c0 + i) + (size_t)(c0 + k) * n] = a[m];
      if (k == i) {  // reciprocal pivot: the substitutions multiply instead of dividing
        const double rp = 1.0 / a[m];
        dg[i] = rp;
        rpiv[c0 + i] = rp;
      }
    }
    __syncthreads();
    return ok;
  }

  __device__ __noinline__ bool factor_diag(int c0, int bs, double* D, double* dg) {
    if (bs == dl::NB) return factor_diag_full(c0, D, dg);
    const int tid = threadIdx.x;
    for (int e = tid; e < bs * bs; e += dl::kThreads) {  // all copies in flight at once
      const int i = e % bs, j = e / bs;
      dl::cp8(D + i + j * dl::DP, K + (c0 + i) + (size_t)(c0 + j) * n, i >= j);
    }
    dl::cp_commit();
    dl::cp_wait<0>();
    __syncthreads();
    bool ok = true;
    for (int j = 0; j < bs; j++) {
      const double d = D[j + j * dl::DP];
      if (!(d > 0.0)) ok = false;
      const double sd = sqrt(d);
      if (tid == 0) dg[j] = sd;
      for (int i = j + 1 + tid; i < bs; i += dl::kThreads) D[i + j * dl::DP] /= sd;
      __syncthreads();
      // trailing (i,k), j < k <= i < bs: thread (tid & 63) owns row i, the
      // eight thread groups stride over the columns
      {
        const int i = j + 1 + (tid & 63);
        if (i < bs) {
          const double lij = D[i + j * dl::DP];
          for (int k = j + 1 + (tid >> 6); k <= i; k += dl::kThreads / 64)
            D[i + k * dl::DP] = fma(-lij, D[k + j * dl::DP], D[i + k * dl::DP]);
        }
      }
      __syncthreads();
    }
    for (int j = tid; j < bs; j += dl::kThreads) {
      D[j + j * dl::DP] = dg[j];
      rpiv[c0 + j] = 1.0 / dg[j];
    }
    __syncthreads();
    for (int e = tid; e < bs * bs; e += dl::kThreads) {
      const int i = e % bs, j = e / bs;
      if (i >= j) K[(c0 + i) + (size_t)(c0 + j) * n] = D[i + j * dl::DP];
    }
    return ok;
  }

  // Rows [r0, n) of block column c0: X = B Lkk^-T on shared-memory tiles of
  // TR rows.  Two threads share a row (even / odd terms of the substitution's
  // dot product, combined with one shuffle); no barrier inside a tile.
  // bs == NB fast path: the tile arrives by cp.async (every copy in flight at
  // once), a thread keeps its half of the row (the x_k with k = half mod 2) in
  // registers through the fully unrolled substitution and stores the result
  // straight to global memory.  Same operation order as the general path.
  __device__ __noinline__ void panel_solve_full(int c0, const double* D, const double* dg,
                                                double* Tm, double* xt_out) {
    constexpr int BS = dl::NB;
    const int tid = threadIdx.x;
    const int half = tid & 1, rl = tid >> 1;
    for (int r0 = c0 + BS; r0 < n; r0 += dl::TR) {
      const int rows = min(dl::TR, n - r0);
      {
        const int r = tid % dl::TR;
        for (int k = tid / dl::TR; k < BS; k += dl::kThreads / dl::TR)
          dl::cp8(Tm + r + k * dl::TR, K + (r0 + min(r, rows - 1)) + (size_t)(c0 + k) * n,
                  r < rows);
        dl::cp_commit();
        dl::cp_wait<0>();
      }
      __syncthreads();
      double x[BS / 2];
#pragma unroll
      for (int m = 0; m < BS / 2; m++) x[m] = Tm[rl + (2 * m + half) * dl::TR];
#pragma unroll
      for (int j = 0; j < BS; j++) {
        // x_j = (b_j - sum_{k<j} x_k L(j,k)) / L(j,j)
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < (j + 1) / 2; m++) {
          const int k = 2 * m + half;
          const double ljk = (2 * m + 1 < j || k < j) ? D[j + k * dl::DP] : 0.0;
          s = fma(x[m], ljk, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        const double xj = (x[j >> 1] - s) * dg[j];  // dg: reciprocal pivots
        if (half == (j & 1)) x[j >> 1] = xj;
      }
      if (rl < rows) {
#pragma unroll
        for (int m = 0; m < BS / 2; m++)
          K[(r0 + rl) + (size_t)(c0 + 2 * m + half) * n] = x[m];
        if (xt_out) {  // row-major copy for the TMA-staged trailing update
#pragma unroll
          for (int m = 0; m < BS / 2; m++) xt_out[(size_t)(r0 + rl) * BS + 2 * m + half] = x[m];
        }
      }
      __syncthreads();
    }
  }

  __device__ __noinline__ void panel_solve(int c0, int bs, const double* D, const double* dg,
                                           double* Tm, double* xt_out) {
    if (bs == dl::NB) {
      panel_solve_full(c0, D, dg, Tm, xt_out);
      return;
    }
    const int tid = threadIdx.x;
    const int half = tid & 1, rl = tid >> 1;
    for (int r0 = c0 + bs; r0 < n; r0 += dl::TR) {
      const int rows = min(dl::TR, n - r0);
      for (int k = tid / dl::TR; k < bs; k += dl::kThreads / dl::TR) {
        const int r = tid % dl::TR;
        if (r < rows) Tm[r + k * dl::TR] = K[(r0 + r) + (size_t)(c0 + k) * n];
      }
      __syncthreads();
      const unsigned mask = __ballot_sync(0xffffffffu, rl < rows);
      if (rl < rows) {
        double* row = Tm + rl;
        for (int j = 0; j < bs; j++) {
          // x_j = (b_j - sum_{k<j} x_k L(j,k)) / L(j,j)
          double s = 0.0;
          for (int k = half; k < j; k += 2) s = fma(row[k * dl::TR], D[j + k * dl::DP], s);
          s += __shfl_xor_sync(mask, s, 1);
          const double xj = (row[j * dl::TR] - s) / D[j + j * dl::DP];
          __syncwarp(mask);
          if (half == 0) row[j * dl::TR] = xj;
          __syncwarp(mask);
        }
      }
      __syncthreads();
      for (int k = tid / dl::TR; k < bs; k += dl::kThreads / dl::TR) {
        const int r = tid % dl::TR;
        if (r < rows) K[(r0 + r) + (size_t)(c0 + k) * n] = Tm[r + k * dl::TR];
      }
      __syncthreads();
    }
  }

  // LinearSolver::Initialize, dense_cholesky_solver.cc:32-79
  __device__ __noinline__ bool factor(const Team& t, const Vars& x, const Vars& xbar,
                         double sigma, double alpha) {
    const int tid = threadIdx.x;
    double* Gam = r2;
    FBS_LAP(15);
    for (int i = tid; i < nv; i += dl::kThreads) {
      const double ys = x.y[i] + sigma * (x.v[i] - xbar.v[i]);
      double ga, mu;
      pfb_barrier(ys, x.v[i], alpha, sigma, &ga, &mu);
      gamma[i] = ga;
      mus[i] = mu;
      Gam[i] = div_nr(ga, mu);
    }
    __syncthreads();
    // E = (H + sigma I) + A' (Gamma A), lower tiles -> K(0:nz, 0:nz)   (:52,62-63)
    {
      const double* Ap = A;
      const int nvv = nv, nzz = nz;
      unsigned seq = 0;
      if (tmA) {
        // the scratch was last touched through the generic proxy (panel tiles,
        // cp.async rings); order those accesses before the TMA engine's writes
        tma::fence_proxy_async();
        __syncthreads();
        seq = hdr->seq;
      }
      for (int I = 0; I * dl::TB < nz; I++)
        for (int J = 0; J * dl::TBN < nz && J * dl::TBN < (I + 1) * dl::TB; J++) {
          double acc[4][4][2];
          auto li = [=](int idx, int k) -> const double* {
            const int c = I * dl::TB + idx;
            return c < nzz ? Ap + k + (size_t)c * nvv : nullptr;
          };
          auto lj = [=](int idx, int k) -> const double* {
            const int c = J * dl::TBN + idx;
            return c < nzz ? Ap + k + (size_t)c * nvv : nullptr;
          };
          const double* Hp = H;
          dl::init_acc(acc, [&](int r, int c) -> double {
            const int gr = I * dl::TB + r, gc = J * dl::TBN + c;
            return (gr < nzz && gc <= gr)
                       ? __ldg(Hp + gr + (size_t)gc * nzz) + (gr == gc ? sigma : 0.0)
                       : 0.0;
          });
          const double one[4] = {1.0, 1.0, 1.0, 1.0};
          const int off = J * dl::TBN - I * dl::TB;  // >= 0: a tile on the diagonal
          const int wm = (tid >> 5) >> 1, wn = (tid >> 5) & 1;
          const bool active = off + 32 * wn <= 32 * wm + 31;
          if (tmA)
            dl::tile_tma<1>(acc, tma::smem_addr(sm), tma::smem_addr(hdr->full), tmA,
                            tma_row0 + I * dl::TB, tma_row0 + J * dl::TBN,
                            off >= 0 ? off : -1, active, Gam, one, nv, seq);
          else
            dl::mma_tile<1>(acc, sm, nv, li, lj, Gam, off >= 0 ? off : -1, active, one);
          dl::for_each_acc(acc, [&](int r, int c, double v) {
            const int gr = I * dl::TB + r, gc = J * dl::TBN + c;
            if (gr < nz && gc <= gr) K[gr + (size_t)gc * n] = v;
          });
        }
      if (tmA) {
        __syncthreads();
        if (tid == 0) hdr->seq = seq;
      }
    }
    FBS_LAP(1);
    // rows of G below E, and S initialised to sigma I   (:67-69 with the sign of
    // the Schur complement folded into the updates)
    for (int e0 = tid; e0 < nl * n; e0 += 8 * dl::kThreads) {  // 8 loads in flight per thread
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + u * dl::kThreads;
        const int r = e % nl, c = e / nl;
        t[u] = (e < nl * n && c < nz) ? __ldg(G + r + (size_t)c * nl) : ((c - nz == r) ? sigma : 0.0);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + u * dl::kThreads;
        const int r = e % nl, c = e / nl;
        if (e < nl * n) K[nz + r + (size_t)c * n] = t[u];
      }
    }
    __syncthreads();
    FBS_LAP(2);
    // blocked right-looking Cholesky over the E block columns, then the S ones
    bool ok = true;
    double* D = sm;
    double* dg = sm + dl::NB * dl::DP;
    double* Tm = sm + dl::kDiag;
    for (int c0 = 0; c0 < n;) {
      const bool inE = c0 < nz;
      const int bs = min(dl::NB, (inE ? nz : n) - c0);
      ok = factor_diag(c0, bs, D, dg) && ok;
      FBS_LAP(3);
      const bool tmaX = tmX != nullptr && bs == dl::NB;
      panel_solve(c0, bs, D, dg, Tm, tmaX ? xt : nullptr);
      FBS_LAP(4);
      unsigned seq = 0;
      if (tmaX) {
        // Xt was written with ordinary stores and the scratch through the generic
        // proxy: order both before the TMA engine's accesses
        tma::fence_proxy_async_all();
        __syncthreads();
        seq = hdr->seq;
      }
      // trailing update with the panel X = K(c0+bs.., c0..c0+bs)
      const int t0 = c0 + bs;
      const double* Kp = K;
      const int nn = n, nzz = nz;
      for (int I = 0; t0 + I * dl::TB < n; I++)
        for (int J = 0; t0 + J * dl::TBN < n && J * dl::TBN < (I + 1) * dl::TB; J++) {
          double acc[4][4][2];
          auto li = [=](int idx, int k) -> const double* {
            const int r = t0 + I * dl::TB + idx;
            return r < nn ? Kp + r + (size_t)(c0 + k) * nn : nullptr;
          };
          auto lj = [=](int idx, int k) -> const double* {
            const int r = t0 + J * dl::TBN + idx;
            return r < nn ? Kp + r + (size_t)(c0 + k) * nn : nullptr;
          };
          dl::init_acc(acc, [&](int r, int c) -> double {
            const int gr = t0 + I * dl::TB + r, gc = t0 + J * dl::TBN + c;
            return (gr < nn && gc <= gr) ? Kp[gr + (size_t)gc * nn] : 0.0;
          });
          // S = sigma I + W'W grows while E's columns are eliminated; everything
          // else shrinks: the sign rides on the B operand's column
          double csgn[4];
          const int wm = (tid >> 5) >> 1, wn = (tid >> 5) & 1;
          {
            const int lane = tid & 31;
#pragma unroll
            for (int b = 0; b < 4; b++) {
              const int gc = t0 + J * dl::TBN + 32 * wn + 8 * b + (lane >> 2);
              csgn[b] = (inE && gc >= nzz) ? 1.0 : -1.0;
            }
          }
          const int off = J * dl::TBN - I * dl::TB;
          const bool active = off + 32 * wn <= 32 * wm + 31;
          if (tmaX)
            dl::tile_tma<2>(acc, tma::smem_addr(sm), tma::smem_addr(hdr->full), tmX,
                            xt_row0 + t0 + I * dl::TB, xt_row0 + t0 + J * dl::TBN,
                            off >= 0 ? off : -1, active, nullptr, csgn, bs, seq);
          else
            dl::mma_tile<2>(acc, sm, bs, li, lj, Kp, off >= 0 ? off : -1, active, csgn);
          dl::for_each_acc(acc, [&](int r, int c, double v) {
            const int gr = t0 + I * dl::TB + r, gc = t0 + J * dl::TBN + c;
            if (gr < n && gc <= gr) K[gr + (size_t)gc * n] = v;
          });
        }
      __syncthreads();
      if (tmaX && tid == 0) hdr->seq = seq;
      FBS_LAP(5);
      c0 += bs;
    }
    return ok;
  }

  // One warp: solves the bs x bs lower-triangular system Lkk u = a (forward)
  // or Lkk' u = a (backward) for the diagonal block at c0; a, u in shared memory.
  // D: the block staged in shared memory (stride DP).
  __device__ __forceinline__ void diag_trsv(const double* D, const double* rp, int bs,
                                            double* a, bool transposed) {
    const int lane = threadIdx.x & 31;
    double v0 = (lane < bs) ? a[lane] : 0.0;
    double v1 = (lane + 32 < bs) ? a[lane + 32] : 0.0;
    if (!transposed) {
      for (int j = 0; j < bs; j++) {
        const double* col = D + j * dl::DP;
        const double aj = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31);
        const double uj = aj * rp[j];
        if (lane == (j & 31)) (j < 32 ? v0 : v1) = uj;
        if (lane > j && lane < bs) v0 = fma(-col[lane], uj, v0);
        if (lane + 32 > j && lane + 32 < bs) v1 = fma(-col[lane + 32], uj, v1);
      }
    } else {
      for (int j = bs - 1; j >= 0; j--) {
        // row j of Lkk' is L(j, i), i < j
        const double aj = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31);
        const double uj = aj * rp[j];
        if (lane == (j & 31)) (j < 32 ? v0 : v1) = uj;
        if (lane < j) v0 = fma(-D[j + lane * dl::DP], uj, v0);
        if (lane + 32 < j) v1 = fma(-D[j + (lane + 32) * dl::DP], uj, v1);
      }
    }
    if (lane < bs) a[lane] = v0;
    if (lane + 32 < bs) a[lane + 32] = v1;
  }

  // LinearSolver::Solve with r = -(rz,rl,rv), dense_cholesky_solver.cc:81-127
  __device__ __noinline__ void solve(const Team& t, const double* rz, const double* rl,
                        const double* rv, const Vars& dx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = dl::kThreads / 32;
    double* ub = sm;           // current block of the right-hand side (NB doubles)
    double* Db = sm + dl::NB;  // its diagonal block of the factor (NB x DP)
    double* rb = Db + dl::NB * dl::DP;  // reciprocal pivots of the block (NB doubles)
    auto stage_block = [&](int c0, int bs) {
      if (tid < bs) {
        ub[tid] = r1[c0 + tid];
        rb[tid] = rpiv[c0 + tid];
      }
      for (int e = tid; e < bs * bs; e += dl::kThreads) {
        const int i = e % bs, j = e / bs;
        if (i >= j) dl::cp8(Db + i + j * dl::DP, K + (c0 + i) + (size_t)(c0 + j) * n, true);
      }
      dl::cp_commit();
      dl::cp_wait<0>();
      __syncthreads();
    };
    FBS_LAP(15);
    for (int i = tid; i < nv; i += dl::kThreads) r2[i] = div_nr(-rv[i], mus[i]);
    for (int i = tid; i < nl; i += dl::kThreads) r1[nz + i] = rl[i];
    __syncthreads();
    for (int i = warp; i < nz; i += nw) {
      // lane's terms k = lane, lane+32, ...: (nv - lane + 31) / 32 of them
      double s = dl::dot_stream(A + (size_t)i * nv + lane, 32, r2 + lane, 32,
                                (nv - lane + 31) / 32, 0.0);
      s = warp_sum(s);
      if (lane == 0) r1[i] = (-rz[i]) - s;
    }
    __syncthreads();
    // forward: u = L^-1 a over the E block columns; the G rows ride along, so
    // r1(nz:n) becomes c - W'u
    auto forward = [&](int cbeg, int cend, int rend) {
      for (int c0 = cbeg; c0 < cend; c0 += dl::NB) {
        const int bs = min(dl::NB, cend - c0);
        stage_block(c0, bs);
        if (warp == 0) diag_trsv(Db, rb, bs, ub, false);
        __syncthreads();
        if (tid < bs) r1[c0 + tid] = ub[tid];
        for (int r = c0 + bs + tid; r < rend; r += dl::kThreads) {
          double s = 0.0;
          for (int k = 0; k < bs; k++) s = fma(K[r + (size_t)(c0 + k) * n], ub[k], s);
          r1[r] -= s;
        }
        __syncthreads();
      }
    };
    // backward: u <- L^-T u over block columns [cbeg, cend)
    auto backward = [&](int cbeg, int cend) {
      int last = cbeg + ((cend - cbeg - 1) / dl::NB) * dl::NB;
      for (int c0 = last; c0 >= cbeg; c0 -= dl::NB) {
        const int bs = min(dl::NB, cend - c0);
        stage_block(c0, bs);
        if (warp == 0) diag_trsv(Db, rb, bs, ub, true);
        __syncthreads();
        if (tid < bs) r1[c0 + tid] = ub[tid];
        // u(j) -= sum_i L(c0+i, j) u(c0+i) for the columns j left of the block
        for (int j0 = cbeg + warp; j0 < c0; j0 += 4 * nw) {  // four columns in flight per warp
          double s[4] = {0.0, 0.0, 0.0, 0.0};
          for (int i = lane; i < bs; i += 32) {
            const double ui = ub[i];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int j = j0 + u * nw;
              if (j < c0) s[u] = fma(K[(size_t)j * n + c0 + i], ui, s[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int j = j0 + u * nw;
            const double su = warp_sum(s[u]);
            if (lane == 0 && j < c0) r1[j] -= su;
          }
        }
        __syncthreads();
      }
    };
    FBS_LAP(6);
    forward(0, nz, n);
    FBS_LAP(7);
    // dl = S^-1 (W'u - c): the tail now holds c - W'u
    for (int i = tid; i < nl; i += dl::kThreads) r1[nz + i] = -r1[nz + i];
    __syncthreads();
    forward(nz, n, n);
    backward(nz, n);
    // u <- u - W dl, W = (K(nz:n, 0:nz))'
    for (int j0 = warp; j0 < nz; j0 += 4 * nw) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = lane; i < nl; i += 32) {
        const double di = r1[nz + i];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u * nw;
          if (j < nz) s[u] = fma(K[(size_t)j * n + nz + i], di, s[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u * nw;
        const double su = warp_sum(s[u]);
        if (lane == 0 && j < nz) r1[j] -= su;
      }
    }
    __syncthreads();
    FBS_LAP(8);
    backward(0, nz);
    FBS_LAP(9);
    for (int i = tid; i < nz; i += dl::kThreads) dx.z[i] = r1[i];
    for (int i = tid; i < nl; i += dl::kThreads) dx.l[i] = r1[nz + i];
    __syncthreads();
    // dv = (rv + gamma .* (A dz)) ./ mus ; dy = b - A dz
    for (int i = tid; i < nv; i += dl::kThreads) {
      const double s = dl::dot_stream(A + i, (size_t)nv, dx.z, 1, nz, 0.0);
      dx.v[i] = div_nr(gamma[i] * s + (-rv[i]), mus[i]);
      dx.y[i] = bvec[i] - s;
    }
    __syncthreads();
    FBS_LAP(10);
  }
};

}  // namespace fbs

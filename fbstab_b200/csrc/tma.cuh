// tma.cuh -- bulk asynchronous copies (the TMA engine's 1-D form), mbarriers
// and the proxy fences that order them against ordinary shared-memory traffic.
//
// Used to stream per-stage matrix blocks global -> shared ahead of the Riccati
// sweep (mpc_riccati.cu) and factor blocks shared -> global behind it.
// cp.async.bulk moves 16-byte aligned runs of whole 16-byte units; a run that is
// not (odd-sized blocks) is read in place instead of being staged.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fbs {
namespace tma {

__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes freshly initialised mbarriers visible to the async proxy
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders generic-proxy shared-memory accesses before subsequent async-proxy ones
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// the same for every state space (global memory written with ordinary stores
// and then read by the TMA engine)
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// global -> shared, bytes % 16 == 0, both addresses 16-byte aligned;
// completion is signalled on `bar` as a transaction count
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes,
                                         unsigned bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// shared -> global, same constraints; tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all committed groups have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// all committed groups are complete (writes visible)
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// A run of n doubles can go through the TMA engine as ONE bulk copy iff it is
// 16-byte aligned and a whole number of 16-byte units.  Runs that are not (in the
// four OCP shapes of the BASELINE configs: only the nu x nu block R, nu odd, whose
// stage-i copy starts at an odd double for every other i) are NOT staged: the
// consumer reads them in place and the issuing thread prefetches their lines.
// (Round 1 moved the ragged first / last double as 8-byte cp.async tracked on the
// TMA mbarrier; compute-sanitizer's racecheck and synccheck do not model that
// protocol and reported it -- profiles/r2_sanitizer.txt.  With whole runs either
// bulk or in place there is no non-bulk asynchronous copy left on the barrier.)
__device__ __forceinline__ bool bulk_able(const double* src, int n) {
  return n > 0 && (((uintptr_t)src >> 3) & 1) == 0 && (n & 1) == 0;
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
// Bytes the run contributes to the stage's transaction count (pass 1, issue = false)
// or issues its bulk copy (pass 2); runs that are not bulk-able are prefetched.
__device__ __forceinline__ unsigned copy_run(double* dst, const double* src, int n,
                                             unsigned bar, bool issue) {
  if (n <= 0) return 0;
  if (!bulk_able(src, n)) {
    if (issue)
      for (int o = 0; o < n; o += 16) prefetch_l1(src + o);
    return 0;
  }
  if (issue) bulk_g2s(smem_addr(dst), src, (unsigned)n * 8u, bar);
  return (unsigned)n * 8u;
}

}  // namespace tma
}  // namespace fbs

// tma.cuh -- bulk asynchronous copies (the TMA engine's 1-D form), mbarriers
// and the proxy fences that order them against ordinary shared-memory traffic.
//
// Used to stream per-stage matrix blocks global -> shared ahead of the Riccati
// sweep (mpc_riccati.cu) and factor blocks shared -> global behind it.
// cp.async.bulk moves 16-byte aligned runs; the at most one leading and one
// trailing double of a run that is only 8-byte aligned travel as 8-byte
// cp.async (LDGSTS) operations tracked by the same mbarrier.
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

// one double, global -> shared, asynchronous (LDGSTS)
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// `bar` receives one arrival when all cp.async of this thread so far are done
// (the pending count is raised by one now and lowered on completion)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// Copies n doubles src -> dst (shared) asynchronously; dst must have the same
// 16-byte phase as src.  Returns the bytes that will be reported to the
// mbarrier as a transaction count (the bulk part); *ragged is set when 8-byte
// cp.async pieces were issued.  Call order for one stage: every copy_run with
// issue_bulk = false first (ragged ends only, sums the bulk bytes), then
// cp_async_mbar_arrive if *ragged, then mbar_arrive_expect_tx(bytes), then
// every copy_run again with issue_bulk = true.
__device__ __forceinline__ unsigned copy_run(double* dst, const double* src, int n,
                                             unsigned bar, bool issue_bulk, bool* ragged) {
  if (n <= 0) return 0;
  const int head = (int)(((uintptr_t)src >> 3) & 1);
  const int interior = (n - head) & ~1;
  const int tail = n - head - interior;
  if (!issue_bulk) {
    if (head) {
      cp_async8(smem_addr(dst), src);
      *ragged = true;
    }
    if (tail) {
      cp_async8(smem_addr(dst + head + interior), src + head + interior);
      *ragged = true;
    }
  } else if (interior) {
    bulk_g2s(smem_addr(dst + head), src + head, (unsigned)interior * 8u, bar);
  }
  return (unsigned)interior * 8u;
}

}  // namespace tma
}  // namespace fbs

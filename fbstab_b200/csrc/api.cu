// api.cu -- C-ABI of the batched engine (include/fbstab_b200.h): handles,
// host<->device staging, kernel selection and launches.
//
// Kernels here are persistent: a fixed grid of CTAs pulls instance indices
// from a global atomic counter, so instances that converge early simply free
// their CTA for the next instance -- no host round trips, no per-iteration
// launches, natural load balance over the 9..28-iteration spread.

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dense_large.cuh"
#include "dense_problem.cuh"
#include "dense_small.h"
#include "engine.cuh"
#include "engine_args.cuh"
#include "fbstab_b200.h"
#include "mpc_lane.h"
#include "mpc_riccati.h"
#include "sparse_lane.h"

namespace {

using fbs::CommonArgs;
using fbs::RunComponent;

thread_local std::string g_last_error;

int Fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess)                                                    \
      return Fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver \
                      ? FBSTAB_ERR_NOGPU                                      \
                      : FBSTAB_ERR_CUDA,                                      \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));        \
  } while (0)

int EnvInt(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// Device (or managed) pointer?  *device receives the owning device of a plain
// device allocation (-1 for managed memory, which every device can read).
bool IsDevicePtr(const void* p, int* device = nullptr) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (device) *device = a.type == cudaMemoryTypeDevice ? a.device : -1;
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Lazily grown device staging buffer for one host-side argument.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int Ensure(size_t bytes) {
    if (bytes <= cap) return FBSTAB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return Fail(FBSTAB_ERR_ALLOC, "cudaMalloc of a staging buffer failed");
    }
    cap = bytes;
    return FBSTAB_OK;
  }
  void Free() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PendingCopy {
  void* host;
  const void* dev;
  size_t bytes;
};
struct RangeCopy {  // deferred mode: `stride` bytes per instance
  void* dev;
  void* host;
  size_t stride;
};

// Resolves user pointers to device pointers, staging host memory.  In deferred
// mode (the batched solves) no copy is issued here: CopyIn / CopyOut move the
// rows of an instance range, so that the H2D copy of chunk c+1 overlaps the
// kernel of chunk c and the D2H copy of chunk c-1.
struct Stager {
  cudaStream_t stream;
  int device = -1;  // the handle's device: device pointers must live there
  bool any_host = false;
  bool defer = false;
  int batch = 1;
  std::vector<PendingCopy> d2h;
  std::vector<RangeCopy> rin, rout;
  int In(DevBuf* buf, const void* user, size_t bytes, const void** out) {
    if (bytes == 0) {
      *out = nullptr;
      return FBSTAB_OK;
    }
    if (!user) return Fail(FBSTAB_ERR_INVALID, "null input pointer");
    int pd = -1;
    if (IsDevicePtr(user, &pd)) {
      if (pd >= 0 && device >= 0 && pd != device)
        return Fail(FBSTAB_ERR_INVALID, "device pointer belongs to another device than the handle");
      *out = user;
      return FBSTAB_OK;
    }
    any_host = true;
    int rc = buf->Ensure(bytes);
    if (rc) return rc;
    if (defer)
      rin.push_back({buf->p, const_cast<void*>(user), bytes / batch});
    else
      CUDA_TRY(cudaMemcpyAsync(buf->p, user, bytes, cudaMemcpyHostToDevice, stream));
    *out = buf->p;
    return FBSTAB_OK;
  }
  // in/out or out-only buffer; copy_in = stage the current host contents
  int InOut(DevBuf* buf, void* user, size_t bytes, bool copy_in, void** out) {
    if (bytes == 0) {
      *out = nullptr;
      return FBSTAB_OK;
    }
    if (!user) return Fail(FBSTAB_ERR_INVALID, "null output pointer");
    int pd = -1;
    if (IsDevicePtr(user, &pd)) {
      if (pd >= 0 && device >= 0 && pd != device)
        return Fail(FBSTAB_ERR_INVALID, "device pointer belongs to another device than the handle");
      *out = user;
      return FBSTAB_OK;
    }
    any_host = true;
    int rc = buf->Ensure(bytes);
    if (rc) return rc;
    if (defer) {
      if (copy_in) rin.push_back({buf->p, user, bytes / batch});
      rout.push_back({buf->p, user, bytes / batch});
    } else {
      if (copy_in)
        CUDA_TRY(cudaMemcpyAsync(buf->p, user, bytes, cudaMemcpyHostToDevice, stream));
      d2h.push_back({user, buf->p, bytes});
    }
    *out = buf->p;
    return FBSTAB_OK;
  }
  int CopyIn(int lo, int n, cudaStream_t s) {
    for (auto& c : rin)
      CUDA_TRY(cudaMemcpyAsync((char*)c.dev + lo * c.stride, (char*)c.host + lo * c.stride,
                               n * c.stride, cudaMemcpyHostToDevice, s));
    return FBSTAB_OK;
  }
  int CopyOut(int lo, int n, cudaStream_t s) {
    for (auto& c : rout)
      CUDA_TRY(cudaMemcpyAsync((char*)c.host + lo * c.stride, (char*)c.dev + lo * c.stride,
                               n * c.stride, cudaMemcpyDeviceToHost, s));
    return FBSTAB_OK;
  }
  int Finish() {
    for (auto& c : d2h)
      CUDA_TRY(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, stream));
    if (any_host) CUDA_TRY(cudaStreamSynchronize(stream));
    return FBSTAB_OK;
  }
};

// ---- options (fbstab_algorithm-impl.h:7-74) --------------------------------
bool Sat(double* x, double a, double b) {
  if (a > b) return false;
  *x = std::max(std::min(*x, b), a);
  return true;
}

struct DenseArgs {
  int nz, nl, nv;
  const double *H, *f, *G, *h, *A, *b;
  CommonArgs c;
};

// dense_large_kernel: the same block plus the batch-wide TMA tensor map of A
struct DenseLargeArgs {
  int nz, nl, nv;
  const double *H, *f, *G, *h, *A, *b;
  CommonArgs c;
  int use_tma;  // 0: cp.async operand staging
  int inst0;    // index of this launch's first instance in the tensor map
  double* xt;   // row-major panel copies, xt_rows x NB doubles per CTA (or nullptr)
  int xt_rows;
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmX;
};

__host__ __device__ inline size_t VecDoubles(int nz, int nl, int nv) {
  return 4 * (size_t)(nz + nl + 2 * nv) + (size_t)(nz + nl + nv);
}

__device__ inline double* Carve(double*& p, size_t n) {
  double* r = p;
  p += n;
  return r;
}

__device__ inline void CarveBuffers(double*& p, int nz, int nl, int nv,
                                    fbs::Buffers* w) {
  fbs::Vars* vs[4] = {&w->xk, &w->xi, &w->xp, &w->dx};
  for (int k = 0; k < 4; k++) {
    vs[k]->z = Carve(p, nz);
    vs[k]->l = Carve(p, nl);
    vs[k]->v = Carve(p, nv);
    vs[k]->y = Carve(p, nv);
  }
  w->ri.z = Carve(p, nz);
  w->ri.l = Carve(p, nl);
  w->ri.v = Carve(p, nv);
}

template <class Args>
__device__ inline void SetupDenseT(const Args& a, int inst, double*& ws,
                                   fbs::DenseProblem* p) {
  p->nz = a.nz;
  p->nl = a.nl;
  p->nv = a.nv;
  p->n = a.nz + a.nl;
  p->H = a.H + (size_t)inst * a.nz * a.nz;
  p->f = a.f + (size_t)inst * a.nz;
  p->G = a.G + (size_t)inst * a.nl * a.nz;
  p->h = a.h + (size_t)inst * a.nl;
  p->A = a.A + (size_t)inst * a.nv * a.nz;
  p->bvec = a.b + (size_t)inst * a.nv;
  p->K = Carve(ws, (size_t)p->n * p->n);
  p->r1 = Carve(ws, p->n);
  p->r2 = Carve(ws, a.nv);
  p->gamma = Carve(ws, a.nv);
  p->mus = Carve(ws, a.nv);
  p->tmp = Carve(ws, p->n);
  p->tz = Carve(ws, a.nz);
}
__host__ __device__ inline size_t DenseDataDoubles(int nz, int nl, int nv) {
  return (size_t)nz * nz + (size_t)nl * nz + (size_t)nv * nz + (size_t)(nz + nl + nv);
}
// Generic dense kernel.  vec_in_smem >= 2: the linear solver's workspace (K and its
// vectors) lives in the CTA's shared memory behind the iterates instead of the per-CTA
// global workspace; == 3: the instance's H, G, A, f, h, b are staged there too, once per
// instance.  Same code, same operation order -- only the address space changes, so the
// results are bit-identical; what changes is that the dependent loads of the left-looking
// LDL' and of the mat-vecs cost a shared-memory round trip instead of an L2 one (the
// reference's own single-QP case, BASELINE config 1, is latency-bound on exactly those).
__device__ inline void SetupDense(const DenseArgs& a, int inst, double*& ws,
                                  fbs::DenseProblem* p) {
  const int mode = a.c.vec_in_smem;
  if (mode < 2) {
    SetupDenseT(a, inst, ws, p);
    return;
  }
  extern __shared__ double dyn_smem[];
  double* sm = dyn_smem + VecDoubles(a.nz, a.nl, a.nv);
  const int nz = a.nz, nl = a.nl, nv = a.nv;
  p->nz = nz;
  p->nl = nl;
  p->nv = nv;
  p->n = nz + nl;
  p->K = Carve(sm, (size_t)p->n * p->n);
  p->r1 = Carve(sm, p->n);
  p->r2 = Carve(sm, nv);
  p->gamma = Carve(sm, nv);
  p->mus = Carve(sm, nv);
  p->tmp = Carve(sm, p->n);
  p->tz = Carve(sm, nz);
  const double* src[6] = {a.H + (size_t)inst * nz * nz, a.f + (size_t)inst * nz,
                          a.G + (size_t)inst * nl * nz, a.h + (size_t)inst * nl,
                          a.A + (size_t)inst * nv * nz, a.b + (size_t)inst * nv};
  if (mode >= 3) {
    const size_t cnt[6] = {(size_t)nz * nz, (size_t)nz, (size_t)nl * nz, (size_t)nl,
                           (size_t)nv * nz, (size_t)nv};
    // (the loop-top barrier of PersistentLoop separates this from the previous instance)
    for (int k = 0; k < 6; k++) {
      double* dst = Carve(sm, cnt[k]);
      for (size_t i = threadIdx.x; i < cnt[k]; i += blockDim.x) dst[i] = __ldg(src[k] + i);
      src[k] = dst;
    }
    __syncthreads();
  }
  p->H = src[0];
  p->f = src[1];
  p->G = src[2];
  p->h = src[3];
  p->A = src[4];
  p->bvec = src[5];
}
size_t DenseWsDoubles(int nz, int nl, int nv) {
  const size_t n = nz + nl;
  return n * n + 2 * n + 3 * (size_t)nv + nz;
}

template <class Args, class P, void (*Setup)(const Args&, int, double*&, P*)>
__device__ void PersistentLoop(const Args& a, int nz, int nl, int nv) {
  extern __shared__ double dyn_smem[];
  __shared__ double red[fbs::kMaxWarps * fbs::kRedSlots];
  __shared__ int s_inst;
  fbs::Team t{red};
  const CommonArgs& c = a.c;
  double* ws0 = c.ws + (size_t)blockIdx.x * c.ws_stride;
#ifdef FBSTAB_PHASE_TIMERS
  if (threadIdx.x == 0) fbs::g_phase_last[blockIdx.x] = clock64();
#endif
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(c.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= c.batch) break;
    double* ws = ws0;
    double* vec = c.vec_in_smem ? dyn_smem : ws;
    fbs::Buffers w;
    CarveBuffers(vec, nz, nl, nv, &w);
    if (!c.vec_in_smem) ws = vec;
    P p;
    Setup(a, inst, ws, &p);
    if (c.comp < 0) {
      fbs::solve_instance(t, p, c.opts, w, c.z + (size_t)inst * nz,
                          c.l + (size_t)inst * nl, c.v + (size_t)inst * nv,
                          c.y + (size_t)inst * nv, c.out + inst);
    } else {
      RunComponent(t, p, c, inst, w);
    }
  }
}

__global__ void __launch_bounds__(256, 2)
dense_generic_kernel(const __grid_constant__ DenseArgs a) {
  PersistentLoop<DenseArgs, fbs::DenseProblem, SetupDense>(a, a.nz, a.nl, a.nv);
}

// Large dense QPs (BASELINE config 5): 256 threads per instance, DMMA SYRK and
// blocked Cholesky (dense_large.cuh); the dynamic shared memory is its scratch.
// Dynamic shared memory of dense_large_kernel: [SmemHeader | pad to 1024 B | scratch]
constexpr size_t kDenseLargeSmem = sizeof(double) * fbs::dl::kSmemDoubles + 2048;
__device__ inline double* DenseLargeScratch(double* dyn) {
  const size_t a = ((size_t)(dyn + 8) + 1023) & ~(size_t)1023;
  return (double*)a;
}
__device__ inline void SetupDenseLarge(const DenseLargeArgs& a, int inst, double*& ws,
                                       fbs::DenseLargeProblem* p) {
  extern __shared__ double dyn_smem[];
  SetupDenseT(a, inst, ws, p);
  p->hdr = (fbs::dl::SmemHeader*)dyn_smem;
  p->sm = DenseLargeScratch(dyn_smem);
  p->rpiv = p->tmp;  // the generic policy's scratch of n doubles
  p->tmA = a.use_tma ? (const void*)&a.tmA : nullptr;
  p->tma_row0 = (a.inst0 + inst) * a.nz;
  if (a.use_tma && a.xt) {
    p->tmX = (const void*)&a.tmX;
    p->xt = a.xt + (size_t)blockIdx.x * a.xt_rows * fbs::dl::NB;
    p->xt_row0 = blockIdx.x * a.xt_rows;
  }
}
__global__ void __launch_bounds__(fbs::dl::kThreads, 2)
dense_large_kernel(const __grid_constant__ DenseLargeArgs a) {
  extern __shared__ double dyn_smem[];
  if (threadIdx.x == 0) {
    fbs::dl::SmemHeader* hdr = (fbs::dl::SmemHeader*)dyn_smem;
    for (int s = 0; s < fbs::dl::kStages; s++)
      fbs::tma::mbar_init(fbs::tma::smem_addr(&hdr->full[s]), 1);
    hdr->seq = 0;
    fbs::tma::fence_mbar_init();
  }
  __syncthreads();
  PersistentLoop<DenseLargeArgs, fbs::DenseLargeProblem, SetupDenseLarge>(a, a.nz, a.nl, a.nv);
}

DenseLargeArgs ToLarge(const DenseArgs& a) {
  DenseLargeArgs r;
  memset(&r, 0, sizeof(r));
  r.nz = a.nz;
  r.nl = a.nl;
  r.nv = a.nv;
  r.H = a.H;
  r.f = a.f;
  r.G = a.G;
  r.h = a.h;
  r.A = a.A;
  r.b = a.b;
  r.c = a.c;
  return r;
}

// Batch-wide tensor map of A (column-major nv x nz per instance, instance-major):
// dim0 = k (nv, contiguous), dim1 = batch * nz columns; box 16 x 64, 128B swizzle.
// Returns false when the shape / alignment rules of the TMA engine are not met
// or the driver entry point is unavailable: the kernel then stages with cp.async.
bool EncodeTensorMapRows(CUtensorMap* tm, const double* base, int inner, long long rows);
bool EncodeTensorMapA(CUtensorMap* tm, const double* A, int nv, int nz, int batch) {
  return EncodeTensorMapRows(tm, A, nv, (long long)batch * nz);
}
// 2-D map of `rows` rows of `inner` contiguous doubles each.
bool EncodeTensorMapRows(CUtensorMap* tm, const double* A, int nv, long long rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                               const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return (EncodeFn) nullptr;
    }
    return (EncodeFn)p;
  }();
  if (!fn || (nv & 1) || nv < fbs::dl::KC || ((size_t)A & 15) || rows > 0x7fffffffLL)
    return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)nv, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)nv * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)fbs::dl::KC, 64};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            EnvInt("FBSTAB_TMA_L2_PROMO", 256) == 256   ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
            : EnvInt("FBSTAB_TMA_L2_PROMO", 256) == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                        : CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- handles ----------------------------------------------------------------
struct HandleBase {
  int device = 0;
  int max_batch = 0;
  int nz = 0, nl = 0, nv = 0;
  fbstab_options opts;
  int sm_count = 0;
  int block = 32;
  int grid_max = 0;
  size_t dyn_smem = 0;
  int vec_in_smem = 0;
  size_t ws_stride = 0;  // doubles
  double* ws = nullptr;
  int* counter = nullptr;
  DevBuf out_buf;
  DevBuf in[12];
  DevBuf io[4];
  DevBuf comp[18];
  int last_launches = 0;
  const char* path = "generic";
  // host-buffer pipeline: copy-in / copy-out streams and per-chunk events
  static constexpr int kMaxChunks = 32;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {};
  // Every launch on a handle shares its instance counter and workspaces, so
  // launches are ordered even when the caller changes streams between calls:
  // each call makes its stream wait for the previous call's last kernel
  // (a no-op when the stream is the same) and records this event behind its own.
  cudaEvent_t ev_last = nullptr;
  bool has_last = false;

  void FreeAll() {
    cudaSetDevice(device);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    for (int i = 0; i < kMaxChunks; i++) {
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_k[i]) cudaEventDestroy(ev_k[i]);
    }
    if (ev_last) cudaEventDestroy(ev_last);
    if (ws) cudaFree(ws);
    if (counter) cudaFree(counter);
    out_buf.Free();
    for (auto& b : in) b.Free();
    for (auto& b : io) b.Free();
    for (auto& b : comp) b.Free();
  }
};

int InitDevice(HandleBase* h, int device, int max_batch) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return Fail(FBSTAB_ERR_NOGPU,
                "no CUDA device available: the engine has no CPU path");
  }
  if (device < 0 || device >= ndev)
    return Fail(FBSTAB_ERR_INVALID, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  h->device = device;
  h->max_batch = max_batch;
  fbstab_default_options(&h->opts);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaMalloc(&h->counter, sizeof(int)));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < HandleBase::kMaxChunks; i++) {
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming));
  return FBSTAB_OK;
}

// Orders this call's work on stream `s` behind the previous call on the handle.
int OrderBegin(HandleBase* h, cudaStream_t s) {
  if (h->has_last) CUDA_TRY(cudaStreamWaitEvent(s, h->ev_last, 0));
  return FBSTAB_OK;
}
int OrderEnd(HandleBase* h, cudaStream_t s) {
  CUDA_TRY(cudaEventRecord(h->ev_last, s));
  h->has_last = true;
  return FBSTAB_OK;
}

// Runs `launch(lo, n)` over the batch.  Device-resident arguments: one launch on
// the caller's stream, asynchronous.  Host buffers: the batch is cut into up to
// kMaxChunks instance ranges of at least `min_chunk` instances (several waves
// of the persistent grid each) and pipelined -- H2D of chunk c+1 on the
// copy-in stream, kernel of chunk c on the caller's stream, D2H of chunk c-1
// on the copy-out stream -- and the call returns when the results are in host
// memory.
template <class Launch>
int RunPipelinedBody(HandleBase* h, Stager* st, int batch, int min_chunk, Launch launch) {
  cudaStream_t cs = st->stream;
  if (!st->any_host) {
    CUDA_TRY(cudaMemsetAsync(h->counter, 0, sizeof(int), cs));
    int rc = launch(0, batch);
    if (rc) return rc;
    h->last_launches = 1;
    return FBSTAB_OK;
  }
  int nchunk = std::max(1, std::min(HandleBase::kMaxChunks, batch / std::max(min_chunk, 1)));
  nchunk = EnvInt("FBSTAB_PIPELINE_CHUNKS", nchunk);
  nchunk = std::max(1, std::min(HandleBase::kMaxChunks, std::min(nchunk, batch)));
  const int per = (batch + nchunk - 1) / nchunk;
  int launches = 0;
  for (int c = 0, lo = 0; lo < batch; c++, lo += per) {
    const int n = std::min(per, batch - lo);
    int rc = st->CopyIn(lo, n, h->s_in);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_in[c], h->s_in));
    CUDA_TRY(cudaStreamWaitEvent(cs, h->ev_in[c], 0));
    CUDA_TRY(cudaMemsetAsync(h->counter, 0, sizeof(int), cs));
    if ((rc = launch(lo, n))) return rc;
    launches++;
    CUDA_TRY(cudaEventRecord(h->ev_k[c], cs));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_k[c], 0));
    if ((rc = st->CopyOut(lo, n, h->s_out))) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_out));
  CUDA_TRY(cudaStreamSynchronize(cs));
  h->last_launches = launches;
  return FBSTAB_OK;
}
template <class Launch>
int RunPipelined(HandleBase* h, Stager* st, int batch, int min_chunk, Launch launch) {
  cudaStream_t cs = st->stream;
  int rc = OrderBegin(h, cs);
  if (rc) return rc;
  if (st->any_host) {
    // the copy-in stream writes the staging buffers the previous call's kernels
    // may still read
    if (h->has_last) CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_last, 0));
  }
  rc = RunPipelinedBody(h, st, batch, min_chunk, launch);
  if (rc && st->any_host) {
    // a mid-pipeline failure: queued copies still target the caller's host buffers
    // and the staging buffers -- drain before reporting
    const std::string keep = g_last_error;
    cudaStreamSynchronize(h->s_in);
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(h->s_out);
    cudaGetLastError();
    g_last_error = keep;
    return rc;
  }
  if (rc) return rc;
  return OrderEnd(h, cs);
}

// scratch_smem > 0: the kernel uses the dynamic shared memory as scratch and
// keeps the iterate vectors in the global workspace.
int InitCommon(HandleBase* h, int device, int max_batch, const void* kernel,
               size_t ws_doubles_no_vec, int block, size_t scratch_smem = 0) {
  int rc = InitDevice(h, device, max_batch);
  if (rc) return rc;
  h->block = scratch_smem ? block : EnvInt("FBSTAB_BLOCK", block);
  const size_t vec_bytes = VecDoubles(h->nz, h->nl, h->nv) * sizeof(double);
  const size_t smem_max = (size_t)EnvInt("FBSTAB_VEC_SMEM_MAX", 12 * 1024);
  h->vec_in_smem = (!scratch_smem && vec_bytes <= smem_max) ? 1 : 0;
  h->dyn_smem = h->vec_in_smem ? vec_bytes : scratch_smem;
  if (h->dyn_smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)h->dyn_smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, h->block,
                                                         h->dyn_smem));
  if (occ < 1) return Fail(FBSTAB_ERR_CUDA, "kernel does not fit on an SM");
  h->ws_stride = ws_doubles_no_vec + (h->vec_in_smem ? 0 : vec_bytes / sizeof(double));
  // bound the workspace for large problems: at most ~2 GiB
  size_t grid = (size_t)h->sm_count * occ;
  const size_t per_cta = std::max<size_t>(h->ws_stride * sizeof(double), 1);
  const size_t cap = ((size_t)2 << 30) / per_cta;
  grid = std::min(grid, std::max<size_t>(cap, (size_t)h->sm_count));
  grid = std::min<size_t>(grid, (size_t)std::max(max_batch, 1));
  h->grid_max = (int)grid;
  if (cudaMalloc(&h->ws, std::max<size_t>(h->ws_stride * grid, 1) * sizeof(double)) !=
      cudaSuccess) {
    cudaGetLastError();
    return Fail(FBSTAB_ERR_ALLOC, "cudaMalloc of the solver workspace failed");
  }
  return FBSTAB_OK;
}

int SetOptions(HandleBase* h, const fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  fbstab_options c = *o;  // UpdateParameters copies field by field, impl:308-332
  int rc = fbstab_validate_options(&c);
  if (rc) return rc;
  h->opts = c;
  return FBSTAB_OK;
}

}  // namespace

struct fbstab_dense_batch : HandleBase {
  fbs::DenseSmallPlan small;
  bool large = false;
  // large path: row-major panel copies for the TMA-staged trailing updates
  double* xt = nullptr;
  int xt_rows = 0;
  bool xt_map = false;
  CUtensorMap tmX;
};
struct fbstab_mpc_batch : HandleBase {
  int N = 0, nx = 0, nu = 0, nc = 0;
  fbs::MpcPlan plan;
  // lane-per-instance path (small stages, large batches)
  double* lane_ws = nullptr;
  double* lane_sdata = nullptr;  // common stage data (shared-data fast path)
  int* lane_mismatch = nullptr;
  int* cta_mismatch = nullptr;  // common-stage-data flag of the CTA path
  int lane_warps = 0;
  int lane_min = 0;
  char lane_name[320];
  // fbstab_mpc_batch_solve_lti: pinned host copy of the horizon (one stage replicated)
  double* lti_host = nullptr;
  size_t lti_cap = 0;
};

struct fbstab_sparse_batch : HandleBase {
  fbs::SparsePattern pat;  // host copy of the symbolic analysis
  fbs::SparseDev dev;      // the same tables on the device
  int* tables = nullptr;   // one allocation behind dev's pointers
  int warps = 0;  // resident instances (workspace columns) of the lane path
  int team_ctas = 0;  // > 0: the warp-per-instance path (sparse_team.cu), this many CTAs
  char name[240];
};

namespace {

int StageComponentIo(HandleBase* h, Stager* st, int batch,
                     const fbstab_component_io* io, fbstab_component_io* dio) {
  const size_t bz = (size_t)batch * h->nz * sizeof(double),
               bl = (size_t)batch * h->nl * sizeof(double),
               bv = (size_t)batch * h->nv * sizeof(double);
  *dio = *io;
  int rc;
  auto in = [&](int slot, const double* user, size_t bytes, const double** out) {
    if (!user) {
      *out = nullptr;
      return FBSTAB_OK;
    }
    const void* p;
    int r = st->In(&h->comp[slot], user, bytes, &p);
    *out = (const double*)p;
    return r;
  };
  auto out = [&](int slot, void* user, size_t bytes, bool copy_in, void** o) {
    if (!user) {
      *o = nullptr;
      return FBSTAB_OK;
    }
    return st->InOut(&h->comp[slot], user, bytes, copy_in, o);
  };
  if ((rc = in(0, io->z, bz, &dio->z))) return rc;
  if ((rc = in(1, io->l, bl, &dio->l))) return rc;
  if ((rc = in(2, io->v, bv, &dio->v))) return rc;
  if ((rc = in(3, io->y, bv, &dio->y))) return rc;
  if ((rc = in(4, io->zbar, bz, &dio->zbar))) return rc;
  if ((rc = in(5, io->lbar, bl, &dio->lbar))) return rc;
  if ((rc = in(6, io->vbar, bv, &dio->vbar))) return rc;
  if ((rc = out(7, io->rz, bz, true, (void**)&dio->rz))) return rc;
  if ((rc = out(8, io->rl, bl, true, (void**)&dio->rl))) return rc;
  if ((rc = out(9, io->rv, bv, true, (void**)&dio->rv))) return rc;
  if ((rc = out(10, io->dz, bz, false, (void**)&dio->dz))) return rc;
  if ((rc = out(11, io->dl, bl, false, (void**)&dio->dl))) return rc;
  if ((rc = out(12, io->dv, bv, false, (void**)&dio->dv))) return rc;
  if ((rc = out(13, io->dy, bv, false, (void**)&dio->dy))) return rc;
  if ((rc = out(14, io->gamma, bv, false, (void**)&dio->gamma))) return rc;
  if ((rc = out(15, io->mus, bv, false, (void**)&dio->mus))) return rc;
  if ((rc = out(16, io->norms, (size_t)batch * 8 * sizeof(double), false,
                (void**)&dio->norms)))
    return rc;
  if ((rc = out(17, io->status, (size_t)batch * sizeof(int32_t), false,
                (void**)&dio->status)))
    return rc;
  return FBSTAB_OK;
}

void FillCommon(HandleBase* h, CommonArgs* c, int batch) {
  c->batch = batch;
  c->ws = h->ws;
  c->ws_stride = h->ws_stride;
  c->counter = h->counter;
  c->vec_in_smem = h->vec_in_smem;
  c->opts = h->opts;
  c->comp = -1;
  memset(&c->io, 0, sizeof(c->io));
}

void StampSolveTime(fbstab_out* out, int batch, double seconds) {
  for (int i = 0; i < batch; i++) out[i].solve_time = seconds;
}

}  // namespace

extern "C" {

const char* fbstab_last_error(void) { return g_last_error.c_str(); }
// (internal: lets the other translation units of the library set the thread's error)
int fbstab_set_last_error_(int code, const char* msg) { return Fail(code, msg ? msg : ""); }

int fbstab_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void fbstab_default_options(fbstab_options* o) {
  o->sigma0 = 1e-8;
  o->sigma_max = 1e-6;
  o->sigma_min = 1e-12;
  o->alpha = 0.95;
  o->beta = 0.75;
  o->eta = 1e-8;
  o->delta = 0.2;
  o->gamma = 0.1;
  o->abs_tol = 1e-6;
  o->rel_tol = 1e-12;
  o->stall_tol = 1e-10;
  o->infeas_tol = 1e-8;
  o->inner_tol_max = 1e-2;
  o->inner_tol_min = 1e-12;
  o->max_newton_iters = 200;
  o->max_prox_iters = 30;
  o->max_inner_iters = 50;
  o->max_linesearch_iters = 20;
  o->check_feasibility = 1;
  o->nonmonotone_linesearch = 1;
  o->display_level = 1;  // Display::FINAL
  o->refine_steps = 0;
  o->regularize_retries = 0;
}

void fbstab_reliable_options(fbstab_options* o) {
  fbstab_default_options(o);
  o->sigma0 = 1e-4;
  o->sigma_max = 1e-2;
  o->sigma_min = 1e-10;
  o->beta = 0.9;
  o->abs_tol = 1e-4;
  o->rel_tol = 1e-6;
  o->max_linesearch_iters = 40;
  o->max_newton_iters = 500;
  o->max_prox_iters = 100;
  o->nonmonotone_linesearch = 0;
}

int fbstab_validate_options(fbstab_options* o) {
  if (!o) return Fail(FBSTAB_ERR_INVALID, "null options");
  bool ok = true;
  o->sigma0 = std::max(o->sigma0, 1e-10);
  ok = ok && Sat(&o->sigma_max, 1e-6, 1e2);
  ok = ok && Sat(&o->sigma_min, 1e-13, 1e-8);
  ok = ok && Sat(&o->sigma0, o->sigma_min, o->sigma_max);
  ok = ok && Sat(&o->alpha, 0.001, 0.999);
  ok = ok && Sat(&o->beta, 0.1, 0.99);
  ok = ok && Sat(&o->eta, 1e-12, 0.499);
  ok = ok && Sat(&o->delta, 0.0001, 0.99);
  ok = ok && Sat(&o->gamma, 0.001, 0.9);
  o->abs_tol = std::max(o->abs_tol, 1e-14);
  o->rel_tol = std::max(o->rel_tol, 0.0);
  o->stall_tol = std::max(o->stall_tol, 1e-14);
  o->infeas_tol = std::max(o->infeas_tol, 1e-14);
  ok = ok && Sat(&o->inner_tol_max, 1e-8, 1e2);
  ok = ok && Sat(&o->inner_tol_min, 1e-14, 1e-2);
  o->max_newton_iters = std::max(o->max_newton_iters, 1);
  o->max_prox_iters = std::max(o->max_prox_iters, 1);
  o->max_inner_iters = std::max(o->max_inner_iters, 1);
  o->max_linesearch_iters = std::max(o->max_linesearch_iters, 1);
  o->refine_steps = std::min(std::max(o->refine_steps, 0), 1);
  o->regularize_retries = std::min(std::max(o->regularize_retries, 0), 8);
  if (!ok) return Fail(FBSTAB_ERR_INVALID, "saturate(): lower bound above upper bound");
  return FBSTAB_OK;
}

// ---- dense -------------------------------------------------------------------
// Shared-memory residency of the generic dense kernel (SetupDense).  Measured on the B200
// (profiles/r2_dense_generic_resident.txt; every mode returns the same bytes): with the
// data and the workspace resident a LONE instance is 10 % faster (BASELINE config 1,
// 50/10/100: 2.43 -> 2.19 ms), but the larger footprint costs occupancy and a full batch
// is 7-12 % slower; the workspace alone (mode 2) is +9 % on 40/0/80 and -30 % on
// 96/16/180.  So: FBSTAB_DENSE_RESIDENT=-1 (default) takes mode 3 when the handle's
// max_batch fits ONE wave of resident CTAs (the latency regime) and it fits shared memory,
// the global workspace otherwise; 0 / 2 / 3 force a mode where it fits (A/B, tests).
static int ConfigureDenseResident(HandleBase* h) {
  int want = EnvInt("FBSTAB_DENSE_RESIDENT", -1);
  const bool automatic = want < 0;
  if (automatic) want = 3;
  if (want < 2) return FBSTAB_OK;
  int smem_optin = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                                  h->device));
  cudaFuncAttributes fa;
  CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)dense_generic_kernel));
  const size_t room = (size_t)smem_optin - fa.sharedSizeBytes;
  const size_t vec = VecDoubles(h->nz, h->nl, h->nv), ws = DenseWsDoubles(h->nz, h->nl, h->nv),
               data = DenseDataDoubles(h->nz, h->nl, h->nv);
  int mode = 0;
  size_t bytes = 0;
  if (want >= 3 && (vec + ws + data) * sizeof(double) <= room) {
    mode = 3;
    bytes = (vec + ws + data) * sizeof(double);
  } else if ((vec + ws) * sizeof(double) <= room) {
    mode = 2;
    bytes = (vec + ws) * sizeof(double);
  }
  if (!mode) return FBSTAB_OK;
  CUDA_TRY(cudaFuncSetAttribute((const void*)dense_generic_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)room));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &occ, (const void*)dense_generic_kernel, h->block, bytes));
  if (occ < 1) return FBSTAB_OK;  // keep the global workspace
  if (automatic && (mode != 3 || h->max_batch > h->sm_count * occ)) return FBSTAB_OK;
  h->vec_in_smem = mode;
  h->dyn_smem = bytes;
  h->grid_max = std::max(1, std::min(h->grid_max, h->sm_count * occ));
  h->path = mode == 3 ? "dense-generic-cta (CTA per instance, data and LDL' workspace "
                        "resident in shared memory)"
                      : "dense-generic-cta (CTA per instance, LDL' workspace in shared "
                        "memory, data from L2)";
  return FBSTAB_OK;
}

int fbstab_dense_batch_create(int nz, int nl, int nv, int max_batch, int device,
                              fbstab_dense_batch** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  // fbstab_dense.cc:19-23
  if (nz <= 0 || nv <= 0 || nl < 0)
    return Fail(FBSTAB_ERR_INVALID,
                "nz and nv must be positive, nl nonnegative");
  if (max_batch < 1) return Fail(FBSTAB_ERR_INVALID, "max_batch must be >= 1");
  auto* h = new fbstab_dense_batch;
  h->nz = nz;
  h->nl = nl;
  h->nv = nv;
  const int n = nz + nl;
  const int block = n <= 8 ? 32 : n <= 48 ? 64 : n <= 128 ? 128 : 256;
  h->large = nz >= EnvInt("FBSTAB_DENSE_LARGE_MIN", 128) && !EnvInt("FBSTAB_FORCE_GENERIC", 0);
  int rc = h->large
               ? InitCommon(h, device, max_batch, (const void*)dense_large_kernel,
                            DenseWsDoubles(nz, nl, nv), fbs::dl::kThreads, kDenseLargeSmem)
               : InitCommon(h, device, max_batch, (const void*)dense_generic_kernel,
                            DenseWsDoubles(nz, nl, nv), block);
  if (rc == FBSTAB_OK && !h->large) rc = ConfigureDenseResident(h);
  if (rc == FBSTAB_OK && !EnvInt("FBSTAB_FORCE_GENERIC", 0))
    rc = fbs::DenseSmallInit(&h->small, nz, nl, nv, h->sm_count, h->counter);
  if (rc) {
    std::string keep = g_last_error;
    h->FreeAll();
    delete h;
    g_last_error = keep;
    return rc;
  }
  if (h->small.enabled) h->path = h->small.name;
  if (h->large && !h->small.enabled && EnvInt("FBSTAB_DENSE_LARGE_TMA", 1)) {
    // one panel copy (rows of K x NB) per resident CTA, rows padded to the tile height
    h->xt_rows = (n + fbs::dl::TB - 1) / fbs::dl::TB * fbs::dl::TB;
    const size_t bytes = (size_t)h->grid_max * h->xt_rows * fbs::dl::NB * sizeof(double);
    if (cudaMalloc(&h->xt, bytes) == cudaSuccess) {
      cudaMemset(h->xt, 0, bytes);
      h->xt_map = EncodeTensorMapRows(&h->tmX, h->xt, fbs::dl::NB,
                                      (long long)h->grid_max * h->xt_rows);
    } else {
      cudaGetLastError();
      h->xt = nullptr;
    }
  }
  if (h->large)
    h->path = "dense-large-cta (256 thr/instance, 2 CTA/SM, DMMA A'GammaA + blocked Cholesky NB=64)";
  *handle = h;
  return FBSTAB_OK;
}

#ifdef FBSTAB_PHASE_TIMERS
// development builds only: read (and optionally clear) the phase cycle counters
extern "C" int fbstab_debug_phase_cycles(unsigned long long* out32, int reset) {
  if (out32) cudaMemcpyFromSymbol(out32, fbs::g_phase_cycles, 32 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[32] = {};
    cudaMemcpyToSymbol(fbs::g_phase_cycles, z, sizeof(z));
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
#endif

int fbstab_dense_batch_destroy(fbstab_dense_batch* h) {
  if (!h) return FBSTAB_OK;
  if (h->xt) {
    cudaSetDevice(h->device);
    cudaFree(h->xt);
  }
  h->FreeAll();
  delete h;
  return FBSTAB_OK;
}

int fbstab_dense_batch_set_options(fbstab_dense_batch* h, const fbstab_options* o) {
  return SetOptions(h, o);
}
int fbstab_dense_batch_get_options(const fbstab_dense_batch* h, fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  *o = h->opts;
  return FBSTAB_OK;
}
int fbstab_dense_batch_last_launches(const fbstab_dense_batch* h) {
  return h ? h->last_launches : 0;
}
const char* fbstab_dense_batch_path(const fbstab_dense_batch* h) {
  return h ? h->path : "";
}

static int DenseStageData(fbstab_dense_batch* h, Stager* st, int batch,
                          const double* H, const double* f, const double* G,
                          const double* hh, const double* A, const double* b,
                          DenseArgs* a) {
  const size_t nz = h->nz, nl = h->nl, nv = h->nv, B = batch, D = sizeof(double);
  a->nz = h->nz;
  a->nl = h->nl;
  a->nv = h->nv;
  int rc;
  if ((rc = st->In(&h->in[0], H, B * nz * nz * D, (const void**)&a->H))) return rc;
  if ((rc = st->In(&h->in[1], f, B * nz * D, (const void**)&a->f))) return rc;
  if ((rc = st->In(&h->in[2], G, B * nl * nz * D, (const void**)&a->G))) return rc;
  if ((rc = st->In(&h->in[3], hh, B * nl * D, (const void**)&a->h))) return rc;
  if ((rc = st->In(&h->in[4], A, B * nv * nz * D, (const void**)&a->A))) return rc;
  if ((rc = st->In(&h->in[5], b, B * nv * D, (const void**)&a->b))) return rc;
  return FBSTAB_OK;
}

int fbstab_dense_batch_solve(fbstab_dense_batch* h, int batch, const double* H,
                             const double* f, const double* G, const double* hh,
                             const double* A, const double* b, double* z,
                             double* l, double* v, double* y, fbstab_out* out,
                             void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  if (!out) return Fail(FBSTAB_ERR_INVALID, "null out pointer");
  h->last_launches = 0;
  if (batch == 0) return FBSTAB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  const auto t0 = std::chrono::steady_clock::now();
  Stager st;
  st.stream = (cudaStream_t)stream;
  st.device = h->device;
  st.defer = true;
  st.batch = batch;
  DenseArgs a;
  int rc = DenseStageData(h, &st, batch, H, f, G, hh, A, b, &a);
  if (rc) return rc;
  const size_t nz = h->nz, nl = h->nl, nv = h->nv, B = batch, D = sizeof(double);
  FillCommon(h, &a.c, batch);
  if ((rc = st.InOut(&h->io[0], z, B * nz * D, true, (void**)&a.c.z))) return rc;
  if ((rc = st.InOut(&h->io[1], l, B * nl * D, true, (void**)&a.c.l))) return rc;
  if ((rc = st.InOut(&h->io[2], v, B * nv * D, true, (void**)&a.c.v))) return rc;
  if ((rc = st.InOut(&h->io[3], y, B * nv * D, false, (void**)&a.c.y))) return rc;
  if ((rc = st.InOut(&h->out_buf, out, B * sizeof(fbstab_out), false,
                     (void**)&a.c.out)))
    return rc;
  // refinement / regularise-and-retry re-use stored factors: the team kernels have them,
  // the warp kernel eliminates with the right-hand side riding along and keeps none
  const bool use_small =
      h->small.enabled && h->opts.refine_steps == 0 && h->opts.regularize_retries == 0;
  // TMA operand staging of the large path: one tensor map over A of the whole batch
  CUtensorMap tmA;
  const bool use_tma = h->large && !use_small && EnvInt("FBSTAB_DENSE_LARGE_TMA", 1) &&
                       EncodeTensorMapA(&tmA, a.A, h->nv, h->nz, batch);
  auto launch = [&](int lo, int n) -> int {
    DenseArgs c = a;
    c.H += (size_t)lo * nz * nz;
    c.f += (size_t)lo * nz;
    c.G += (size_t)lo * nl * nz;
    c.h += (size_t)lo * nl;
    c.A += (size_t)lo * nv * nz;
    c.b += (size_t)lo * nv;
    c.c.z += (size_t)lo * nz;
    c.c.l += (size_t)lo * nl;
    c.c.v += (size_t)lo * nv;
    c.c.y += (size_t)lo * nv;
    c.c.out += lo;
    c.c.batch = n;
    if (use_small) {
      if (fbs::DenseSmallLaunch(h->small, n, c.H, c.f, c.G, c.h, c.A, c.b, c.c.z, c.c.l,
                                c.c.v, c.c.y, c.c.out, h->opts, -1, nullptr, st.stream))
        return Fail(FBSTAB_ERR_CUDA, "dense small-path launch failed");
    } else {
      const int grid = std::min(n, h->grid_max);
      if (h->large) {
        DenseLargeArgs cl = ToLarge(c);
        cl.A = a.A;  // instances are addressed from the start of the batch (inst0)
        cl.A += (size_t)lo * nv * nz;
        cl.use_tma = use_tma ? 1 : 0;
        cl.inst0 = lo;
        if (use_tma) cl.tmA = tmA;
        if (use_tma && h->xt_map) {
          cl.xt = h->xt;
          cl.xt_rows = h->xt_rows;
          cl.tmX = h->tmX;
        }
        dense_large_kernel<<<grid, h->block, h->dyn_smem, st.stream>>>(cl);
      } else {
        dense_generic_kernel<<<grid, h->block, h->dyn_smem, st.stream>>>(c);
      }
    }
    CUDA_TRY(cudaGetLastError());
    return FBSTAB_OK;
  };
  const int capacity = use_small ? h->small.grid * fbs::DenseSmallWarpsPerCta() : h->grid_max;
  if ((rc = RunPipelined(h, &st, batch, 4 * capacity, launch))) return rc;
  if (st.any_host && !IsDevicePtr(out)) {
    const double sec =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    StampSolveTime(out, batch, sec);
  }
  return FBSTAB_OK;
}

int fbstab_dense_batch_component(fbstab_dense_batch* h, int comp, int batch,
                                 const double* H, const double* f,
                                 const double* G, const double* hh,
                                 const double* A, const double* b,
                                 const fbstab_component_io* io, void* stream) {
  if (!h || !io) return Fail(FBSTAB_ERR_INVALID, "null argument");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  if (comp < FBSTAB_COMP_MARGIN || comp > FBSTAB_COMP_FEAS)
    return Fail(FBSTAB_ERR_INVALID, "unknown component");
  if (batch == 0) return FBSTAB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  Stager st;
  st.stream = (cudaStream_t)stream;
  st.device = h->device;
  DenseArgs a;
  int rc = DenseStageData(h, &st, batch, H, f, G, hh, A, b, &a);
  if (rc) return rc;
  FillCommon(h, &a.c, batch);
  a.c.comp = comp;
  if ((rc = StageComponentIo(h, &st, batch, io, &a.c.io))) return rc;
  if ((rc = OrderBegin(h, st.stream))) return rc;
  CUDA_TRY(cudaMemsetAsync(h->counter, 0, sizeof(int), st.stream));
  if (h->small.enabled) {
    rc = fbs::DenseSmallLaunch(h->small, batch, a.H, a.f, a.G, a.h, a.A, a.b,
                               nullptr, nullptr, nullptr, nullptr, nullptr,
                               h->opts, comp, &a.c.io, st.stream);
    if (rc) return Fail(FBSTAB_ERR_CUDA, "dense small-path launch failed");
  } else {
    const int grid = std::min(batch, h->grid_max);
    if (h->large)
      dense_large_kernel<<<grid, h->block, h->dyn_smem, st.stream>>>(ToLarge(a));
    else
      dense_generic_kernel<<<grid, h->block, h->dyn_smem, st.stream>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  h->last_launches = 1;
  if ((rc = OrderEnd(h, st.stream))) return rc;
  return st.Finish();
}

// ---- mpc ---------------------------------------------------------------------
int fbstab_mpc_batch_create(int N, int nx, int nu, int nc, int max_batch,
                            int device, fbstab_mpc_batch** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  // fbstab_mpc.cc:62-65
  if (N < 1 || nx < 1 || nu < 1 || nc < 1)
    return Fail(FBSTAB_ERR_INVALID, "problem sizes must be positive");
  if (max_batch < 1) return Fail(FBSTAB_ERR_INVALID, "max_batch must be >= 1");
  auto* h = new fbstab_mpc_batch;
  h->N = N;
  h->nx = nx;
  h->nu = nu;
  h->nc = nc;
  h->nz = (N + 1) * (nx + nu);
  h->nl = (N + 1) * nx;
  h->nv = (N + 1) * nc;
  int rc = InitDevice(h, device, max_batch);
  if (rc == FBSTAB_OK) {
    const char* err = "";
    rc = fbs::MpcPlanInit(&h->plan, N, nx, nu, nc, max_batch, h->sm_count, &err);
    if (rc) Fail(rc, err);
  }
  if (rc) {
    std::string keep = g_last_error;
    fbs::MpcPlanFree(&h->plan);
    h->FreeAll();
    delete h;
    g_last_error = keep;
    return rc;
  }
  h->path = h->plan.name;
  if (EnvInt("FBSTAB_MPC_SHARED", 1) && cudaMalloc(&h->cta_mismatch, sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    h->cta_mismatch = nullptr;
  }
  // Small stages: one LANE per instance once the batch is several waves of the
  // CTA kernel (below that the CTA kernel's shorter per-instance latency wins).
  h->lane_min = EnvInt("FBSTAB_MPC_LANE_MIN",
                       std::max(256, (int)(fbs::MpcLaneCrossover(nx, nu, nc) * 16 * h->sm_count)));
  if (fbs::MpcLaneSupported(nx, nu, nc) && EnvInt("FBSTAB_MPC_LANE", 1) &&
      max_batch >= h->lane_min) {
    int warps = fbs::MpcLaneWarps(N, nx, nu, nc, max_batch, h->sm_count, nullptr);
    if (EnvInt("FBSTAB_MPC_LANE_WARPS_PER_SM", 0) > 0)
      warps = std::min(warps, h->sm_count * EnvInt("FBSTAB_MPC_LANE_WARPS_PER_SM", 0));
    const size_t bytes = (size_t)warps * fbs::MpcLaneWsDoublesPerWarp(N, nx, nu, nc) * 8;
    if (cudaMalloc(&h->lane_ws, bytes) == cudaSuccess) {
      h->lane_warps = warps;
      if (EnvInt("FBSTAB_MPC_SHARED", 1)) {
        if (cudaMalloc(&h->lane_sdata, fbs::MpcLaneSharedDoubles(N, nx, nu, nc) * 8) !=
                cudaSuccess ||
            cudaMalloc(&h->lane_mismatch, sizeof(int)) != cudaSuccess) {
          cudaGetLastError();
          if (h->lane_sdata) cudaFree(h->lane_sdata);
          h->lane_sdata = nullptr;
          h->lane_mismatch = nullptr;
        }
      }
      snprintf(h->lane_name, sizeof(h->lane_name),
               "mpc-lane<nx,nu,nc> (lane per instance, %d warps, %.0f MB interleaved "
               "workspace%s; batches < %d: %s)",
               warps, bytes / 1e6,
               h->lane_sdata ? ", common stage data detected on the device and shared" : "",
               h->lane_min, "mpc-riccati-cta");
      h->path = h->lane_name;
    } else {
      cudaGetLastError();
      h->lane_ws = nullptr;
    }
  }
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_mpc_batch_destroy(fbstab_mpc_batch* h) {
  if (!h) return FBSTAB_OK;
  cudaSetDevice(h->device);
  fbs::MpcPlanFree(&h->plan);
  if (h->lane_ws) cudaFree(h->lane_ws);
  if (h->lane_sdata) cudaFree(h->lane_sdata);
  if (h->lane_mismatch) cudaFree(h->lane_mismatch);
  if (h->cta_mismatch) cudaFree(h->cta_mismatch);
  if (h->lti_host) cudaFreeHost(h->lti_host);
  h->FreeAll();
  delete h;
  return FBSTAB_OK;
}
int fbstab_mpc_batch_set_options(fbstab_mpc_batch* h, const fbstab_options* o) {
  return SetOptions(h, o);
}
int fbstab_mpc_batch_get_options(const fbstab_mpc_batch* h, fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  *o = h->opts;
  return FBSTAB_OK;
}
int fbstab_mpc_batch_last_launches(const fbstab_mpc_batch* h) {
  return h ? h->last_launches : 0;
}
const char* fbstab_mpc_batch_path(const fbstab_mpc_batch* h) {
  return h ? h->path : "";
}

// shared: the 11 stage-data sequences are ONE copy for the whole batch (x0 stays per
// instance).  Host copies of shared data are staged at once, ahead of the pipeline.
static int MpcStageData(fbstab_mpc_batch* h, Stager* st, int batch,
                        const double* const* user, fbs::MpcData* a, bool shared = false) {
  const size_t N = h->N, nx = h->nx, nu = h->nu, nc = h->nc, K = N + 1,
               B = batch, D = sizeof(double);
  const size_t sizes[12] = {K * nx * nx, K * nu * nu, K * nu * nx, K * nx,
                            K * nu,      N * nx * nx, N * nx * nu, N * nx,
                            K * nc * nx, K * nc * nu, K * nc,      nx};
  const double** dst[12] = {&a->Q, &a->R, &a->S, &a->q, &a->r, &a->A,
                            &a->B, &a->c, &a->E, &a->L, &a->d, &a->x0};
  for (int k = 0; k < 12; k++) {
    int rc;
    if (shared && k < 11) {
      const bool defer = st->defer;
      st->defer = false;  // immediate copy on the caller's stream
      rc = st->In(&h->in[k], user[k], sizes[k] * D, (const void**)dst[k]);
      st->defer = defer;
    } else {
      rc = st->In(&h->in[k], user[k], B * sizes[k] * D, (const void**)dst[k]);
    }
    if (rc) return rc;
  }
  return FBSTAB_OK;
}

static int MpcSolveImpl(fbstab_mpc_batch* h, int batch, const double* const* user, bool shared,
                        double* z, double* l, double* v, double* y, fbstab_out* out,
                        void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  if (!out) return Fail(FBSTAB_ERR_INVALID, "null out pointer");
  h->last_launches = 0;
  if (batch == 0) return FBSTAB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  const auto t0 = std::chrono::steady_clock::now();
  Stager st;
  st.stream = (cudaStream_t)stream;
  st.device = h->device;
  st.defer = true;
  st.batch = batch;
  fbs::MpcData a;
  int rc;
  if (shared && h->has_last) {
    // the one-copy staging buffers may still be read by the previous call's kernels
    CUDA_TRY(cudaStreamWaitEvent(st.stream, h->ev_last, 0));
  }
  if ((rc = MpcStageData(h, &st, batch, user, &a, shared))) return rc;
  const size_t nz = h->nz, nl = h->nl, nv = h->nv, Bn = batch, D = sizeof(double);
  double *dz, *dl, *dv, *dy;
  fbstab_out* dout;
  if ((rc = st.InOut(&h->io[0], z, Bn * nz * D, true, (void**)&dz))) return rc;
  if ((rc = st.InOut(&h->io[1], l, Bn * nl * D, true, (void**)&dl))) return rc;
  if ((rc = st.InOut(&h->io[2], v, Bn * nv * D, true, (void**)&dv))) return rc;
  if ((rc = st.InOut(&h->io[3], y, Bn * nv * D, false, (void**)&dy))) return rc;
  if ((rc = st.InOut(&h->out_buf, out, Bn * sizeof(fbstab_out), false, (void**)&dout)))
    return rc;
  // shared stage data needs the kernels' common-data mode (a flag in device memory)
  if (shared && !h->cta_mismatch)
    return Fail(FBSTAB_ERR_INVALID, "shared stage data is disabled (FBSTAB_MPC_SHARED=0)");
  auto launch = [&](int lo, int n) -> int {
    const size_t N = h->N, nx = h->nx, nu = h->nu, nc = h->nc, K = N + 1, o = lo;
    fbs::MpcData c = a;
    if (!shared) {
      c.Q += o * K * nx * nx;
      c.R += o * K * nu * nu;
      c.S += o * K * nu * nx;
      c.q += o * K * nx;
      c.r += o * K * nu;
      c.A += o * N * nx * nx;
      c.B += o * N * nx * nu;
      c.c += o * N * nx;
      c.E += o * K * nc * nx;
      c.L += o * K * nc * nu;
      c.d += o * K * nc;
    }
    c.x0 += o * nx;
    const bool lane = h->lane_ws && n >= h->lane_min && (!shared || h->lane_sdata) &&
                      h->opts.refine_steps == 0 && h->opts.regularize_retries == 0;
    int fail;
    if (lane) {
      fail = fbs::MpcLaneLaunch(h->N, h->nx, h->nu, h->nc, n, h->lane_warps, c, dz + o * nz,
                                dl + o * nl, dv + o * nv, dy + o * nv, dout + lo, h->opts,
                                h->lane_ws, h->counter, h->lane_mismatch, h->lane_sdata,
                                st.stream, shared);
    } else if (shared) {
      // explicit common data: the flag is set without the detection pass
      fail = cudaMemsetAsync(h->cta_mismatch, 0, sizeof(int), st.stream) != cudaSuccess ||
             fbs::MpcLaunch(h->plan, n, c, dz + o * nz, dl + o * nl, dv + o * nv, dy + o * nv,
                            dout + lo, h->opts, -1, nullptr, h->counter, h->cta_mismatch,
                            st.stream);
    } else {
      fail = (h->cta_mismatch && n > 1 &&
              fbs::MpcSharedDetect(h->N, h->nx, h->nu, h->nc, n, c, h->cta_mismatch,
                                   st.stream)) ||
             fbs::MpcLaunch(h->plan, n, c, dz + o * nz, dl + o * nl, dv + o * nv, dy + o * nv,
                            dout + lo, h->opts, -1, nullptr, h->counter,
                            n > 1 ? h->cta_mismatch : nullptr, st.stream);
    }
    if (fail) return Fail(FBSTAB_ERR_CUDA, "MPC kernel launch failed");
    CUDA_TRY(cudaGetLastError());
    return FBSTAB_OK;
  };
  if ((rc = RunPipelined(h, &st, batch, h->lane_ws ? 8 * 32 * h->lane_warps : 8 * h->plan.grid_max,
                         launch)))
    return rc;
  if (st.any_host && !IsDevicePtr(out)) {
    const double sec =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    StampSolveTime(out, batch, sec);
  }
  return FBSTAB_OK;
}

int fbstab_mpc_batch_solve(fbstab_mpc_batch* h, int batch, const double* Q,
                           const double* R, const double* S, const double* q,
                           const double* r, const double* A, const double* B,
                           const double* c, const double* E, const double* L,
                           const double* d, const double* x0, double* z,
                           double* l, double* v, double* y, fbstab_out* out,
                           void* stream) {
  const double* user[12] = {Q, R, S, q, r, A, B, c, E, L, d, x0};
  return MpcSolveImpl(h, batch, user, false, z, l, v, y, out, stream);
}

int fbstab_mpc_batch_solve_shared(fbstab_mpc_batch* h, int batch, const double* Q,
                                  const double* R, const double* S, const double* q,
                                  const double* r, const double* A, const double* B,
                                  const double* c, const double* E, const double* L,
                                  const double* d, const double* x0, double* z, double* l,
                                  double* v, double* y, fbstab_out* out, void* stream) {
  const double* user[12] = {Q, R, S, q, r, A, B, c, E, L, d, x0};
  return MpcSolveImpl(h, batch, user, true, z, l, v, y, out, stream);
}

// Time-invariant stage data: ONE stage of each sequence, replicated over the
// horizon exactly like OcpGenerator::CopyOverHorizon (ocp_generator.cc:397-418:
// the same Q,R,S,q,r,L,d at stages 0..N, A,B,c at stages 0..N-1, and E(0) = 0 --
// no constraint on the measured state), then solved as shared stage data.
int fbstab_mpc_batch_solve_lti(fbstab_mpc_batch* h, int batch, const double* Q,
                               const double* R, const double* S, const double* q,
                               const double* r, const double* A, const double* B,
                               const double* c, const double* E, const double* L,
                               const double* d, const double* x0, double* z, double* l,
                               double* v, double* y, fbstab_out* out, void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  const size_t N = h->N, nx = h->nx, nu = h->nu, nc = h->nc, K = N + 1;
  const size_t one[11] = {nx * nx, nu * nu, nu * nx, nx, nu, nx * nx, nx * nu, nx,
                          nc * nx, nc * nu, nc};
  const size_t len[11] = {K, K, K, K, K, N, N, N, K, K, K};
  const double* src[11] = {Q, R, S, q, r, A, B, c, E, L, d};
  size_t total = 0;
  for (int k = 0; k < 11; k++) total += one[k] * len[k];
  CUDA_TRY(cudaSetDevice(h->device));
  // the horizon copy is 11 small sequences: built on the host (pinned, owned by the
  // handle) and handed to the shared-data entry, which stages it with one copy each
  if (h->lti_cap < total) {
    if (h->lti_host) cudaFreeHost(h->lti_host);
    h->lti_host = nullptr;
    h->lti_cap = 0;
    if (cudaMallocHost(&h->lti_host, total * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      return Fail(FBSTAB_ERR_ALLOC, "cudaMallocHost of the LTI horizon copy failed");
    }
    h->lti_cap = total;
  }
  // a previous LTI call's asynchronous staging copies must have left the buffer
  if (h->has_last) CUDA_TRY(cudaEventSynchronize(h->ev_last));
  std::vector<double> tmp;
  const double* user[12];
  double* w = h->lti_host;
  for (int k = 0; k < 11; k++) {
    if (!src[k]) return Fail(FBSTAB_ERR_INVALID, "null input pointer");
    const double* s1 = src[k];
    if (IsDevicePtr(s1)) {  // one stage from the device: a few hundred bytes
      tmp.resize(one[k]);
      CUDA_TRY(cudaMemcpy(tmp.data(), s1, one[k] * sizeof(double), cudaMemcpyDeviceToHost));
      s1 = tmp.data();
    }
    for (size_t i = 0; i < len[k]; i++) {
      if (k == 8 && i == 0)
        memset(w + i * one[k], 0, one[k] * sizeof(double));  // E(0) = 0
      else
        memcpy(w + i * one[k], s1, one[k] * sizeof(double));
    }
    user[k] = w;
    w += one[k] * len[k];
  }
  user[11] = x0;
  return MpcSolveImpl(h, batch, user, true, z, l, v, y, out, stream);
}

int fbstab_mpc_batch_component(fbstab_mpc_batch* h, int comp, int batch,
                               const double* Q, const double* R,
                               const double* S, const double* q,
                               const double* r, const double* A,
                               const double* B, const double* c,
                               const double* E, const double* L,
                               const double* d, const double* x0,
                               const fbstab_component_io* io, void* stream) {
  if (!h || !io) return Fail(FBSTAB_ERR_INVALID, "null argument");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  if (comp < FBSTAB_COMP_MARGIN || comp > FBSTAB_COMP_FEAS)
    return Fail(FBSTAB_ERR_INVALID, "unknown component");
  if (batch == 0) return FBSTAB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  Stager st;
  st.stream = (cudaStream_t)stream;
  st.device = h->device;
  fbs::MpcData a;
  const double* user[12] = {Q, R, S, q, r, A, B, c, E, L, d, x0};
  int rc = MpcStageData(h, &st, batch, user, &a);
  if (rc) return rc;
  fbstab_component_io dio;
  if ((rc = StageComponentIo(h, &st, batch, io, &dio))) return rc;
  if ((rc = OrderBegin(h, st.stream))) return rc;
  CUDA_TRY(cudaMemsetAsync(h->counter, 0, sizeof(int), st.stream));
  if (fbs::MpcLaunch(h->plan, batch, a, nullptr, nullptr, nullptr, nullptr, nullptr,
                     h->opts, comp, &dio, h->counter, nullptr, st.stream))
    return Fail(FBSTAB_ERR_CUDA, "MPC kernel launch failed");
  CUDA_TRY(cudaGetLastError());
  h->last_launches = 1;
  if ((rc = OrderEnd(h, st.stream))) return rc;
  return st.Finish();
}

/* ---- sparse QPs with a common pattern (FBstabSparse) ------------------------------- */
int fbstab_sparse_batch_create(int nz, int nl, int nv, const int* Hp, const int* Hi,
                               const int* Gp, const int* Gi, const int* Ap, const int* Ai,
                               const int* perm, int max_batch, int device,
                               fbstab_sparse_batch** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  if (max_batch < 1) return Fail(FBSTAB_ERR_INVALID, "max_batch must be >= 1");
  auto* h = new fbstab_sparse_batch();
  if (!fbs::SparseAnalyze(nz, nl, nv, Hp, Hi, Gp, Gi, Ap, Ai, perm, &h->pat)) {
    const std::string msg = "sparse pattern: " + h->pat.error;
    delete h;
    return Fail(FBSTAB_ERR_INVALID, msg);
  }
  h->nz = nz;
  h->nl = nl;
  h->nv = nv;
  int rc = InitDevice(h, device, max_batch);
  auto fail = [&](int code) {
    const std::string keep = g_last_error;
    if (h->tables) cudaFree(h->tables);
    h->FreeAll();
    delete h;
    g_last_error = keep;
    return code;
  };
  if (rc) return fail(rc);
  const fbs::SparsePattern& p = h->pat;
  // all integer tables in one device allocation
  const std::vector<int>* tabs[] = {&p.Hr_ptr, &p.Hr_col, &p.Hr_val, &p.Gp,  &p.Gi,    &p.Gr_ptr,
                                    &p.Gr_col, &p.Gr_val, &p.Ap,     &p.Ai,  &p.Ar_ptr, &p.Ar_col,
                                    &p.Ar_val, &p.iperm,  &p.Kp,     &p.Ki,  &p.Kkind,  &p.Kidx,
                                    &p.Krow,   &p.Lp,     &p.Li,     &p.Sp,  &p.Sc,     &p.St};
  constexpr int kTabs = sizeof(tabs) / sizeof(tabs[0]);
  size_t off[kTabs + 1];
  off[0] = 0;
  for (int k = 0; k < kTabs; k++) off[k + 1] = off[k] + ((tabs[k]->size() + 3) & ~(size_t)3) + 4;
  if (cudaMalloc(&h->tables, off[kTabs] * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    Fail(FBSTAB_ERR_ALLOC, "cudaMalloc of the sparse pattern tables failed");
    return fail(FBSTAB_ERR_ALLOC);
  }
  for (int k = 0; k < kTabs; k++)
    if (!tabs[k]->empty() &&
        cudaMemcpy(h->tables + off[k], tabs[k]->data(), tabs[k]->size() * sizeof(int),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
      Fail(FBSTAB_ERR_CUDA, "copy of the sparse pattern tables failed");
      return fail(FBSTAB_ERR_CUDA);
    }
  fbs::SparseDev& d = h->dev;
  d.nz = nz; d.nl = nl; d.nv = nv; d.n = p.n;
  d.nnzH = p.nnzH; d.nnzG = p.nnzG; d.nnzA = p.nnzA; d.nnzK = p.nnzK; d.nnzL = p.nnzL;
  const int** dst[] = {&d.Hr_ptr, &d.Hr_col, &d.Hr_val, &d.Gp,  &d.Gi,    &d.Gr_ptr,
                       &d.Gr_col, &d.Gr_val, &d.Ap,     &d.Ai,  &d.Ar_ptr, &d.Ar_col,
                       &d.Ar_val, &d.iperm,  &d.Kp,     &d.Ki,  &d.Kkind,  &d.Kidx,
                       &d.Krow,   &d.Lp,     &d.Li,     &d.Sp,  &d.Sc,     &d.St};
  for (int k = 0; k < kTabs; k++) *dst[k] = h->tables + off[k];
  // Which device path: the warp-per-instance kernel's throughput is its residency over
  // its per-instance latency whatever the batch (cfg 3a as sparse QPs: 37 k solves/s);
  // the lane-per-instance kernel needs tens of thousands of instances in flight to
  // cover its latency (14 k at 16,384, 37 k at 65,536) -- it takes the large batches.
  const int team_occ = fbs::SparseTeamCtasPerSm(d);
  if (team_occ > 0 && max_batch < EnvInt("FBSTAB_SPARSE_TEAM_MAX_BATCH", 98304)) {
    // warp per instance: the factor and the LDL' work vectors live in shared memory
    const int ctas = std::min(max_batch, h->sm_count * team_occ);
    const size_t per_cta = fbs::SparseTeamWsDoubles(d) * sizeof(double);
    if (cudaMalloc(&h->ws, (size_t)ctas * per_cta) != cudaSuccess) {
      cudaGetLastError();
      Fail(FBSTAB_ERR_ALLOC, "cudaMalloc of the sparse solver workspace failed");
      return fail(FBSTAB_ERR_ALLOC);
    }
    h->team_ctas = ctas;
    snprintf(h->name, sizeof(h->name),
             "sparse-team (warp per instance, common pattern: n=%d nnz(K)=%d nnz(L)=%d, L and "
             "the LDL' work vectors in %.1f KB of shared memory, %d CTA/SM)",
             p.n, p.nnzK, p.nnzL, fbs::SparseTeamSmemBytes(d) / 1024.0, team_occ);
  } else {
  // lane per instance: one workspace column per resident instance, at most ~6 GiB
  const size_t per_lane = fbs::SparseLaneWsDoublesPerLane(d) * sizeof(double);
  size_t slots = fbs::SparseLaneSlots(max_batch, h->sm_count);
  slots = std::max<size_t>(32, std::min<size_t>(slots, ((size_t)6 << 30) / std::max<size_t>(per_lane, 1)));
  if (cudaMalloc(&h->ws, slots * per_lane) != cudaSuccess) {
    cudaGetLastError();
    Fail(FBSTAB_ERR_ALLOC, "cudaMalloc of the sparse solver workspace failed");
    return fail(FBSTAB_ERR_ALLOC);
  }
  h->warps = (int)slots;
  snprintf(h->name, sizeof(h->name),
           "sparse-lane (lane per instance, common pattern: n=%d nnz(K)=%d nnz(L)=%d, "
           "%d resident instances, %d per warp, %.0f MB interleaved workspace)",
           p.n, p.nnzK, p.nnzL, (int)slots, fbs::SparseLaneLanes(max_batch, h->sm_count),
           slots * per_lane / 1e6);
  }
  h->path = h->name;
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_sparse_batch_destroy(fbstab_sparse_batch* h) {
  if (!h) return FBSTAB_OK;
  cudaSetDevice(h->device);
  if (h->tables) cudaFree(h->tables);
  h->FreeAll();
  delete h;
  return FBSTAB_OK;
}
int fbstab_sparse_batch_set_options(fbstab_sparse_batch* h, const fbstab_options* o) {
  return SetOptions(h, o);
}
int fbstab_sparse_batch_get_options(const fbstab_sparse_batch* h, fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  *o = h->opts;
  return FBSTAB_OK;
}
int fbstab_sparse_batch_last_launches(const fbstab_sparse_batch* h) {
  return h ? h->last_launches : 0;
}
const char* fbstab_sparse_batch_path(const fbstab_sparse_batch* h) { return h ? h->path : ""; }
int fbstab_sparse_batch_analysis(const fbstab_sparse_batch* h, int* n, int* nnzK, int* nnzL,
                                 int* perm) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (n) *n = h->pat.n;
  if (nnzK) *nnzK = h->pat.nnzK;
  if (nnzL) *nnzL = h->pat.nnzL;
  if (perm) std::copy(h->pat.perm.begin(), h->pat.perm.end(), perm);
  return FBSTAB_OK;
}

// Host only (no device is touched): the symbolic analysis fbstab_sparse_batch_create
// would run for this pattern.
int fbstab_sparse_analyze(int nz, int nl, int nv, const int* Hp, const int* Hi, const int* Gp,
                          const int* Gi, const int* Ap, const int* Ai, const int* user_perm,
                          int* n, int* nnzK, int* nnzL, int* perm) {
  if (nz <= 0 || nl < 0 || nv <= 0 || !Hp || !Hi || !Ap || !Ai || (nl > 0 && !Gp))
    return Fail(FBSTAB_ERR_INVALID, "In FBstabSparse::FBstabSparse: invalid sizes or null pattern");
  fbs::SparsePattern pat;
  if (!fbs::SparseAnalyze(nz, nl, nv, Hp, Hi, Gp, Gi, Ap, Ai, user_perm, &pat))
    return Fail(FBSTAB_ERR_INVALID, pat.error.c_str());
  if (n) *n = pat.n;
  if (nnzK) *nnzK = pat.nnzK;
  if (nnzL) *nnzL = pat.nnzL;
  if (perm) std::copy(pat.perm.begin(), pat.perm.end(), perm);
  return FBSTAB_OK;
}

int fbstab_sparse_batch_factor_pattern(const fbstab_sparse_batch* h, int* Lp, int* Li) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (Lp) std::copy(h->pat.Lp.begin(), h->pat.Lp.end(), Lp);
  if (Li) std::copy(h->pat.Li.begin(), h->pat.Li.end(), Li);
  return FBSTAB_OK;
}

int fbstab_sparse_batch_solve(fbstab_sparse_batch* h, int batch, const double* Hx,
                              const double* f, const double* Gx, const double* hh,
                              const double* Ax, const double* b, double* z, double* l,
                              double* v, double* y, fbstab_out* out, void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  if (!out) return Fail(FBSTAB_ERR_INVALID, "null out pointer");
  if (h->team_ctas == 0 && (h->opts.refine_steps != 0 || h->opts.regularize_retries != 0))
    return Fail(FBSTAB_ERR_INVALID,
                "refine_steps / regularize_retries: not on the lane-per-instance sparse path");
  h->last_launches = 0;
  if (batch == 0) return FBSTAB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  const auto t0 = std::chrono::steady_clock::now();
  Stager st;
  st.stream = (cudaStream_t)stream;
  st.device = h->device;
  st.defer = true;
  st.batch = batch;
  const size_t nz = h->nz, nl = h->nl, nv = h->nv, B = batch, D = sizeof(double);
  const size_t nH = h->pat.nnzH, nG = h->pat.nnzG, nA = h->pat.nnzA;
  const double *dH, *df, *dG, *dh, *dA, *db;
  double *dz, *dl, *dv, *dy;
  fbstab_out* dout;
  int rc;
  if ((rc = st.In(&h->in[0], Hx, B * nH * D, (const void**)&dH))) return rc;
  if ((rc = st.In(&h->in[1], f, B * nz * D, (const void**)&df))) return rc;
  if ((rc = st.In(&h->in[2], Gx, B * nG * D, (const void**)&dG))) return rc;
  if ((rc = st.In(&h->in[3], hh, B * nl * D, (const void**)&dh))) return rc;
  if ((rc = st.In(&h->in[4], Ax, B * nA * D, (const void**)&dA))) return rc;
  if ((rc = st.In(&h->in[5], b, B * nv * D, (const void**)&db))) return rc;
  if ((rc = st.InOut(&h->io[0], z, B * nz * D, true, (void**)&dz))) return rc;
  if ((rc = st.InOut(&h->io[1], l, B * nl * D, true, (void**)&dl))) return rc;
  if ((rc = st.InOut(&h->io[2], v, B * nv * D, true, (void**)&dv))) return rc;
  if ((rc = st.InOut(&h->io[3], y, B * nv * D, false, (void**)&dy))) return rc;
  if ((rc = st.InOut(&h->out_buf, out, B * sizeof(fbstab_out), false, (void**)&dout))) return rc;
  auto launch = [&](int lo, int n) -> int {
    const size_t o = (size_t)lo;
    const int rc2 =
        h->team_ctas > 0
            ? fbs::SparseTeamLaunch(h->dev, n, h->team_ctas, dH + o * nH, df + o * nz,
                                    dG + o * nG, dh + o * nl, dA + o * nA, db + o * nv,
                                    dz + o * nz, dl + o * nl, dv + o * nv, dy + o * nv, dout + lo,
                                    h->opts, h->ws, h->counter, st.stream)
            : fbs::SparseLaneLaunch(h->dev, n, h->warps, dH + o * nH, df + o * nz, dG + o * nG,
                                    dh + o * nl, dA + o * nA, db + o * nv, dz + o * nz,
                                    dl + o * nl, dv + o * nv, dy + o * nv, dout + lo, h->opts,
                                    h->ws, h->counter, st.stream);
    if (rc2) return Fail(FBSTAB_ERR_CUDA, "sparse kernel launch failed");
    return FBSTAB_OK;
  };
  if ((rc = RunPipelined(h, &st, batch, 8 * std::max(h->warps, h->team_ctas), launch))) return rc;
  if (st.any_host && !IsDevicePtr(out)) {
    const double sec =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    StampSolveTime(out, batch, sec);
  }
  return FBSTAB_OK;
}

}  // extern "C"


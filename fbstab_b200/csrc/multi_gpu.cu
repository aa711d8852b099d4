// multi_gpu.cu -- the batch across the GPUs of one box (include/fbstab_b200.h,
// "multi-GPU").
//
// QP instances are independent (reference fbstab/fbstab_dense.h:137-142: one
// Solve = one problem), so the batch shards by contiguous instance ranges and
// the solve itself needs NO collective.  Two host-side drivers:
//
//  * one process per GPU (torchrun, MPI, ...): every rank owns a communicator
//    (fbstab_multi_gpu_create) and solves its shard with its own batch handle;
//    the only exchange is fbstab_multi_gpu_gather -- one grouped NCCL
//    send/recv per result array, straight from each rank's result buffers into
//    the root's global arrays over NVLink (no packing, no padding, no staging
//    copy: a shard's rows are already contiguous in instance-major layout);
//
//  * one process driving several GPUs (fbstab_*_multi_gpu_solve, what the C++
//    facade's SolveBatch(..., devices) calls): one host thread per device runs
//    the ordinary pipelined host-buffer solve on its shard; results land in the
//    caller's host arrays at the shard's offset, so there is nothing to gather.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a process that already
// carries a NCCL -- PyTorch does -- shares it instead of loading a second copy,
// and the library still loads on a box without NCCL (the entry points then
// return FBSTAB_ERR_NCCL).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "fbstab_b200.h"

namespace {

// ---- the slice of nccl.h this file uses (ABI stable across NCCL 2.x) ---------
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
enum { kNcclChar = 0, kNcclFloat64 = 8 };

struct Nccl {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const Nccl& GetNccl() {
  static Nccl n = [] {
    Nccl r;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      r.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (r.so) break;
    }
    if (!r.so) return r;
    auto sym = [&](const char* s) { return dlsym(r.so, s); };
    r.GetUniqueId = (decltype(r.GetUniqueId))sym("ncclGetUniqueId");
    r.CommInitRank = (decltype(r.CommInitRank))sym("ncclCommInitRank");
    r.CommDestroy = (decltype(r.CommDestroy))sym("ncclCommDestroy");
    r.GroupStart = (decltype(r.GroupStart))sym("ncclGroupStart");
    r.GroupEnd = (decltype(r.GroupEnd))sym("ncclGroupEnd");
    r.Send = (decltype(r.Send))sym("ncclSend");
    r.Recv = (decltype(r.Recv))sym("ncclRecv");
    r.GetErrorString = (decltype(r.GetErrorString))sym("ncclGetErrorString");
    r.ok = r.GetUniqueId && r.CommInitRank && r.CommDestroy && r.GroupStart && r.GroupEnd &&
           r.Send && r.Recv;
    return r;
  }();
  return n;
}

}  // namespace

// api.cu owns the thread-local error string
extern "C" int fbstab_set_last_error_(int code, const char* msg);

namespace {

int Fail(int code, const std::string& msg) { return fbstab_set_last_error_(code, msg.c_str()); }

int NcclFail(ncclResult_t r, const char* what) {
  const Nccl& n = GetNccl();
  return Fail(FBSTAB_ERR_NCCL, std::string(what) + ": " +
                                   (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
}

void ShardRange(int nranks, int rank, long total, long* first, long* count) {
  // contiguous ranges of ceil / floor size: the first (total % nranks) ranks take one more
  const long base = total / nranks, rem = total % nranks;
  *first = rank * base + std::min<long>(rank, rem);
  *count = base + (rank < rem ? 1 : 0);
}

}  // namespace

struct fbstab_multi_gpu {
  int rank = 0, nranks = 1, device = 0;
  ncclComm_t comm = nullptr;
};

extern "C" {

int fbstab_multi_gpu_shard(int nranks, int rank, long global_batch, long* first, long* count) {
  if (nranks < 1 || rank < 0 || rank >= nranks || global_batch < 0 || !first || !count)
    return Fail(FBSTAB_ERR_INVALID, "fbstab_multi_gpu_shard: bad arguments");
  ShardRange(nranks, rank, global_batch, first, count);
  return FBSTAB_OK;
}

int fbstab_multi_gpu_unique_id(char id[128]) {
  const Nccl& n = GetNccl();
  if (!n.ok) return Fail(FBSTAB_ERR_NCCL, "libnccl.so.2 is not available");
  if (!id) return Fail(FBSTAB_ERR_INVALID, "null id");
  ncclUniqueId u;
  ncclResult_t r = n.GetUniqueId(&u);
  if (r) return NcclFail(r, "ncclGetUniqueId");
  memcpy(id, u.internal, 128);
  return FBSTAB_OK;
}

int fbstab_multi_gpu_create(int rank, int nranks, const char id[128], int device,
                            fbstab_multi_gpu** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks)
    return Fail(FBSTAB_ERR_INVALID, "rank / nranks out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return Fail(FBSTAB_ERR_NOGPU, "no CUDA device available: the engine has no CPU path");
  }
  if (device < 0 || device >= ndev) return Fail(FBSTAB_ERR_INVALID, "device index out of range");
  auto* h = new fbstab_multi_gpu;
  h->rank = rank;
  h->nranks = nranks;
  h->device = device;
  if (nranks > 1) {
    const Nccl& n = GetNccl();
    if (!n.ok) {
      delete h;
      return Fail(FBSTAB_ERR_NCCL, "libnccl.so.2 is not available");
    }
    if (!id) {
      delete h;
      return Fail(FBSTAB_ERR_INVALID, "null id");
    }
    if (cudaSetDevice(device) != cudaSuccess) {
      cudaGetLastError();
      delete h;
      return Fail(FBSTAB_ERR_CUDA, "cudaSetDevice failed");
    }
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclResult_t r = n.CommInitRank(&h->comm, nranks, u, rank);
    if (r) {
      delete h;
      return NcclFail(r, "ncclCommInitRank");
    }
  }
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_multi_gpu_destroy(fbstab_multi_gpu* h) {
  if (!h) return FBSTAB_OK;
  if (h->comm) {
    cudaSetDevice(h->device);
    GetNccl().CommDestroy(h->comm);
  }
  delete h;
  return FBSTAB_OK;
}

int fbstab_multi_gpu_gather(fbstab_multi_gpu* h, int root, long global_batch, int nz, int nl,
                            int nv, const double* z, const double* l, const double* v,
                            const double* y, const fbstab_out* out, double* Z, double* L,
                            double* V, double* Y, fbstab_out* OUT, void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (root < 0 || root >= h->nranks || global_batch < 0 || nz < 1 || nl < 0 || nv < 1)
    return Fail(FBSTAB_ERR_INVALID, "fbstab_multi_gpu_gather: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaSetDevice(h->device) != cudaSuccess) {
    cudaGetLastError();
    return Fail(FBSTAB_ERR_CUDA, "cudaSetDevice failed");
  }
  long first, count;
  ShardRange(h->nranks, h->rank, global_batch, &first, &count);
  // the five result arrays: (source, destination, bytes per instance)
  struct Row {
    const void* src;
    void* dst;
    size_t width;
  };
  const Row rows[5] = {{z, Z, sizeof(double) * (size_t)nz},
                       {l, L, sizeof(double) * (size_t)nl},
                       {v, V, sizeof(double) * (size_t)nv},
                       {y, Y, sizeof(double) * (size_t)nv},
                       {out, OUT, sizeof(fbstab_out)}};
  const bool is_root = h->rank == root;
  for (const Row& r : rows) {
    if (r.width == 0) continue;
    if ((count > 0 && !r.src) || (is_root && !r.dst))
      return Fail(FBSTAB_ERR_INVALID, "fbstab_multi_gpu_gather: null result array");
  }
  if (is_root) {
    // the root's own rows: a device-to-device copy (no-op when the shard already
    // lives inside the global arrays)
    for (const Row& r : rows) {
      char* dst = (char*)r.dst + (size_t)first * r.width;
      if (r.width && count > 0 && dst != r.src &&
          cudaMemcpyAsync(dst, r.src, (size_t)count * r.width, cudaMemcpyDeviceToDevice, s) !=
              cudaSuccess)
        return Fail(FBSTAB_ERR_CUDA, "device copy of the root's shard failed");
    }
  }
  if (h->nranks == 1) return FBSTAB_OK;
  const Nccl& n = GetNccl();
  ncclResult_t rc = n.GroupStart();
  if (rc) return NcclFail(rc, "ncclGroupStart");
  for (const Row& r : rows) {
    if (r.width == 0) continue;
    if (is_root) {
      for (int p = 0; p < h->nranks && !rc; p++) {
        if (p == root) continue;
        long pf, pc;
        ShardRange(h->nranks, p, global_batch, &pf, &pc);
        if (pc > 0)
          rc = n.Recv((char*)r.dst + (size_t)pf * r.width, (size_t)pc * r.width, kNcclChar, p,
                      h->comm, s);
      }
    } else if (count > 0) {
      rc = n.Send(r.src, (size_t)count * r.width, kNcclChar, root, h->comm, s);
    }
    if (rc) break;
  }
  ncclResult_t re = n.GroupEnd();
  if (rc) return NcclFail(rc, "ncclSend / ncclRecv");
  if (re) return NcclFail(re, "ncclGroupEnd");
  return FBSTAB_OK;
}

// ---- one process, several GPUs: host buffers in, host buffers out ------------------
struct fbstab_dense_multi_gpu {
  int nz = 0, nl = 0, nv = 0;
  std::vector<int> devices;
  std::vector<fbstab_dense_batch*> shards;
  long max_batch = 0;
};
struct fbstab_mpc_multi_gpu {
  int N = 0, nx = 0, nu = 0, nc = 0;
  std::vector<int> devices;
  std::vector<fbstab_mpc_batch*> shards;
  long max_batch = 0;
};

struct fbstab_sparse_multi_gpu {
  int nz = 0, nl = 0, nv = 0;
  size_t nnzH = 0, nnzG = 0, nnzA = 0;
  std::vector<int> devices;
  std::vector<fbstab_sparse_batch*> shards;
  long max_batch = 0;
};

static int CheckDevices(int ndev, const int* devices) {
  if (ndev < 1 || !devices) return Fail(FBSTAB_ERR_INVALID, "empty device list");
  const int have = fbstab_device_count();
  if (have == 0)
    return Fail(FBSTAB_ERR_NOGPU, "no CUDA device available: the engine has no CPU path");
  for (int i = 0; i < ndev; i++) {
    if (devices[i] < 0 || devices[i] >= have)
      return Fail(FBSTAB_ERR_INVALID, "device index out of range");
    for (int j = 0; j < i; j++)
      if (devices[j] == devices[i]) return Fail(FBSTAB_ERR_INVALID, "duplicate device");
  }
  return FBSTAB_OK;
}

int fbstab_dense_multi_gpu_create(int ndev, const int* devices, int nz, int nl, int nv,
                                  long max_batch, fbstab_dense_multi_gpu** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  int rc = CheckDevices(ndev, devices);
  if (rc) return rc;
  if (max_batch < 1) return Fail(FBSTAB_ERR_INVALID, "max_batch must be >= 1");
  auto* h = new fbstab_dense_multi_gpu;
  h->nz = nz;
  h->nl = nl;
  h->nv = nv;
  h->max_batch = max_batch;
  h->devices.assign(devices, devices + ndev);
  for (int i = 0; i < ndev; i++) {
    long first, count;
    ShardRange(ndev, i, max_batch, &first, &count);
    fbstab_dense_batch* s = nullptr;
    rc = fbstab_dense_batch_create(nz, nl, nv, (int)std::max<long>(count, 1), devices[i], &s);
    if (rc) {
      for (auto* p : h->shards) fbstab_dense_batch_destroy(p);
      delete h;
      return rc;
    }
    h->shards.push_back(s);
  }
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_dense_multi_gpu_destroy(fbstab_dense_multi_gpu* h) {
  if (!h) return FBSTAB_OK;
  for (auto* p : h->shards) fbstab_dense_batch_destroy(p);
  delete h;
  return FBSTAB_OK;
}

int fbstab_dense_multi_gpu_set_options(fbstab_dense_multi_gpu* h, const fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  for (auto* p : h->shards) {
    int rc = fbstab_dense_batch_set_options(p, o);
    if (rc) return rc;
  }
  return FBSTAB_OK;
}

int fbstab_dense_multi_gpu_solve(fbstab_dense_multi_gpu* h, long batch, const double* H,
                                 const double* f, const double* G, const double* hh,
                                 const double* A, const double* b, double* z, double* l,
                                 double* v, double* y, fbstab_out* out) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  const int nd = (int)h->shards.size();
  const size_t nz = h->nz, nl = h->nl, nv = h->nv;
  std::vector<int> rcs(nd, FBSTAB_OK);
  std::vector<std::string> errs(nd);
  std::vector<std::thread> th;
  for (int i = 0; i < nd; i++) {
    long first, count;
    ShardRange(nd, i, batch, &first, &count);
    if (count == 0) continue;
    th.emplace_back([=, &rcs, &errs] {
      const size_t o = (size_t)first;
      rcs[i] = fbstab_dense_batch_solve(
          h->shards[i], (int)count, H + o * nz * nz, f + o * nz, G ? G + o * nl * nz : G,
          hh ? hh + o * nl : hh, A + o * nv * nz, b + o * nv, z + o * nz, l ? l + o * nl : l,
          v + o * nv, y + o * nv, out + o, nullptr);
      if (rcs[i]) errs[i] = fbstab_last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < nd; i++)
    if (rcs[i]) return Fail(rcs[i], "device " + std::to_string(h->devices[i]) + ": " + errs[i]);
  return FBSTAB_OK;
}

int fbstab_mpc_multi_gpu_create(int ndev, const int* devices, int N, int nx, int nu, int nc,
                                long max_batch, fbstab_mpc_multi_gpu** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  int rc = CheckDevices(ndev, devices);
  if (rc) return rc;
  if (max_batch < 1) return Fail(FBSTAB_ERR_INVALID, "max_batch must be >= 1");
  auto* h = new fbstab_mpc_multi_gpu;
  h->N = N;
  h->nx = nx;
  h->nu = nu;
  h->nc = nc;
  h->max_batch = max_batch;
  h->devices.assign(devices, devices + ndev);
  for (int i = 0; i < ndev; i++) {
    long first, count;
    ShardRange(ndev, i, max_batch, &first, &count);
    fbstab_mpc_batch* s = nullptr;
    rc = fbstab_mpc_batch_create(N, nx, nu, nc, (int)std::max<long>(count, 1), devices[i], &s);
    if (rc) {
      for (auto* p : h->shards) fbstab_mpc_batch_destroy(p);
      delete h;
      return rc;
    }
    h->shards.push_back(s);
  }
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_mpc_multi_gpu_destroy(fbstab_mpc_multi_gpu* h) {
  if (!h) return FBSTAB_OK;
  for (auto* p : h->shards) fbstab_mpc_batch_destroy(p);
  delete h;
  return FBSTAB_OK;
}

int fbstab_mpc_multi_gpu_set_options(fbstab_mpc_multi_gpu* h, const fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  for (auto* p : h->shards) {
    int rc = fbstab_mpc_batch_set_options(p, o);
    if (rc) return rc;
  }
  return FBSTAB_OK;
}

int fbstab_mpc_multi_gpu_solve(fbstab_mpc_multi_gpu* h, long batch, const double* Q,
                               const double* R, const double* S, const double* q,
                               const double* r, const double* A, const double* B,
                               const double* c, const double* E, const double* L,
                               const double* d, const double* x0, double* z, double* l,
                               double* v, double* y, fbstab_out* out) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  const int nd = (int)h->shards.size();
  const size_t N = h->N, nx = h->nx, nu = h->nu, nc = h->nc, K = N + 1;
  const size_t nz = K * (nx + nu), nl = K * nx, nv = K * nc;
  std::vector<int> rcs(nd, FBSTAB_OK);
  std::vector<std::string> errs(nd);
  std::vector<std::thread> th;
  for (int i = 0; i < nd; i++) {
    long first, count;
    ShardRange(nd, i, batch, &first, &count);
    if (count == 0) continue;
    th.emplace_back([=, &rcs, &errs] {
      const size_t o = (size_t)first;
      rcs[i] = fbstab_mpc_batch_solve(
          h->shards[i], (int)count, Q + o * K * nx * nx, R + o * K * nu * nu,
          S + o * K * nu * nx, q + o * K * nx, r + o * K * nu, A + o * N * nx * nx,
          B + o * N * nx * nu, c + o * N * nx, E + o * K * nc * nx, L + o * K * nc * nu,
          d + o * K * nc, x0 + o * nx, z + o * nz, l + o * nl, v + o * nv, y + o * nv, out + o,
          nullptr);
      if (rcs[i]) errs[i] = fbstab_last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < nd; i++)
    if (rcs[i]) return Fail(rcs[i], "device " + std::to_string(h->devices[i]) + ": " + errs[i]);
  return FBSTAB_OK;
}

int fbstab_sparse_multi_gpu_create(int ndev, const int* devices, int nz, int nl, int nv,
                                   const int* Hp, const int* Hi, const int* Gp, const int* Gi,
                                   const int* Ap, const int* Ai, const int* perm,
                                   long max_batch, fbstab_sparse_multi_gpu** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  int rc = CheckDevices(ndev, devices);
  if (rc) return rc;
  if (max_batch < 1) return Fail(FBSTAB_ERR_INVALID, "max_batch must be >= 1");
  if (nz <= 0 || !Hp || !Ap || (nl > 0 && !Gp))
    return Fail(FBSTAB_ERR_INVALID, "In FBstabSparse::FBstabSparse: invalid sizes or null pattern");
  auto* h = new fbstab_sparse_multi_gpu;
  h->nz = nz;
  h->nl = nl;
  h->nv = nv;
  h->nnzH = (size_t)Hp[nz];
  h->nnzG = nl > 0 ? (size_t)Gp[nz] : 0;
  h->nnzA = (size_t)Ap[nz];
  h->max_batch = max_batch;
  h->devices.assign(devices, devices + ndev);
  for (int i = 0; i < ndev; i++) {
    long first, count;
    ShardRange(ndev, i, max_batch, &first, &count);
    fbstab_sparse_batch* s = nullptr;
    // every shard runs the same symbolic analysis: same elimination order on every device
    rc = fbstab_sparse_batch_create(nz, nl, nv, Hp, Hi, Gp, Gi, Ap, Ai, perm,
                                    (int)std::max<long>(count, 1), devices[i], &s);
    if (rc) {
      for (auto* p : h->shards) fbstab_sparse_batch_destroy(p);
      delete h;
      return rc;
    }
    h->shards.push_back(s);
  }
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_sparse_multi_gpu_destroy(fbstab_sparse_multi_gpu* h) {
  if (!h) return FBSTAB_OK;
  for (auto* p : h->shards) fbstab_sparse_batch_destroy(p);
  delete h;
  return FBSTAB_OK;
}

int fbstab_sparse_multi_gpu_set_options(fbstab_sparse_multi_gpu* h, const fbstab_options* o) {
  if (!h || !o) return Fail(FBSTAB_ERR_INVALID, "null argument");
  for (auto* p : h->shards) {
    int rc = fbstab_sparse_batch_set_options(p, o);
    if (rc) return rc;
  }
  return FBSTAB_OK;
}

int fbstab_sparse_multi_gpu_solve(fbstab_sparse_multi_gpu* h, long batch, const double* Hx,
                                  const double* f, const double* Gx, const double* hh,
                                  const double* Ax, const double* b, double* z, double* l,
                                  double* v, double* y, fbstab_out* out) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (batch < 0 || batch > h->max_batch)
    return Fail(FBSTAB_ERR_INVALID, "batch exceeds the handle's max_batch");
  const int nd = (int)h->shards.size();
  const size_t nz = h->nz, nl = h->nl, nv = h->nv;
  const size_t nH = h->nnzH, nG = h->nnzG, nA = h->nnzA;
  std::vector<int> rcs(nd, FBSTAB_OK);
  std::vector<std::string> errs(nd);
  std::vector<std::thread> th;
  for (int i = 0; i < nd; i++) {
    long first, count;
    ShardRange(nd, i, batch, &first, &count);
    if (count == 0) continue;
    th.emplace_back([=, &rcs, &errs] {
      const size_t o = (size_t)first;
      rcs[i] = fbstab_sparse_batch_solve(
          h->shards[i], (int)count, Hx + o * nH, f + o * nz, Gx ? Gx + o * nG : Gx,
          hh ? hh + o * nl : hh, Ax + o * nA, b + o * nv, z + o * nz, l ? l + o * nl : l,
          v + o * nv, y + o * nv, out + o, nullptr);
      if (rcs[i]) errs[i] = fbstab_last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < nd; i++)
    if (rcs[i]) return Fail(rcs[i], "device " + std::to_string(h->devices[i]) + ": " + errs[i]);
  return FBSTAB_OK;
}

}  // extern "C"

// closed_loop.cu -- batched receding-horizon (closed-loop) MPC as one library
// object (include/fbstab_b200.h, "closed loop").
//
// This is what the reference's OcpGenerator::GetSimulationInputs exists for
// (fbstab/test/ocp_generator.h:31-38,69; ocp_generator.cc:56-71: the plant
// matrices A, B, C, D, the initial state and the number of steps T) and what its
// README means by "can be easily warmstarted" (README.md:20).  B plants are
// simulated for T control steps; one step is, for every plant at once,
//
//   1. warm start: shift the previous solution by one stage (stage i <- stage
//      i + 1, the last stage repeated) -- or start cold;
//   2. solve the OCP from the measured state x(t)   (x0 of the wire format);
//   3. apply the first input u(t) = u_0 to the plant, x(t+1) = A x(t) + B u(t) + c,
//      and log x, u.
//
// Everything stays on the device between steps: the handle owns the problem data,
// the iterates, the states and the logs; a step enqueues three kernels (shift,
// the persistent solve kernel, plant update) on the caller's stream and returns.
// A plant whose OCP ends with an infeasibility flag or a failed factorisation
// does not get the certificate / stale iterate applied as an input: its previous
// input is held and its warm start is reset (the next solve starts cold).
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "fbstab_b200.h"

extern "C" int fbstab_set_last_error_(int code, const char* msg);

namespace {

int Fail(int code, const std::string& msg) { return fbstab_set_last_error_(code, msg.c_str()); }

#define CL_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return Fail(FBSTAB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// stage-major iterate of K stages, w doubles per stage: stage i <- stage i+1.
// One thread per (instance, entry of a stage) walks the horizon front to back, so
// every element is read before it is overwritten.
__global__ void shift_kernel(double* z, double* l, double* v, int batch, int K, int wz, int wl,
                             int wv) {
  const int per = wz + wl + wv;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)batch * per) return;
  const int inst = (int)(t / per), e = (int)(t % per);
  double* a;
  int w, o;
  if (e < wz) {
    a = z + (size_t)inst * K * wz, w = wz, o = e;
  } else if (e < wz + wl) {
    a = l + (size_t)inst * K * wl, w = wl, o = e - wz;
  } else {
    a = v + (size_t)inst * K * wv, w = wv, o = e - wz - wl;
  }
  for (int i = 0; i + 1 < K; i++) a[(size_t)i * w + o] = a[(size_t)(i + 1) * w + o];
}

// One thread per plant: u = first input of the solution (or the held input),
// x <- A x + B u + c, logs.  A, B, c: per plant (stride > 0) or common (stride 0),
// column-major.  hold: plants whose solve did not end in SUCCESS / MAXITERATIONS.
__global__ void plant_kernel(const double* A, const double* B, const double* c, size_t sA,
                             size_t sB, size_t sc, int batch, int nx, int nu, int K, double* x,
                             double* u_prev, double* z, double* l, double* v, int wl, int wv,
                             const fbstab_out* out, double* X, double* U, int step, int steps) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= batch) return;
  const int w = nx + nu;
  double* zi = z + (size_t)p * K * w;
  const bool good = out[p].status == FBSTAB_STATUS_OK &&
                    (out[p].eflag == FBSTAB_SUCCESS || out[p].eflag == FBSTAB_MAXITERATIONS);
  double* up = u_prev + (size_t)p * nu;
  if (good) {
    for (int j = 0; j < nu; j++) up[j] = zi[nx + j];
  } else {
    // the iterate holds a certificate or a stale point: next solve starts cold
    for (int e = 0; e < K * w; e++) zi[e] = 0.0;
    for (int e = 0; e < K * wl; e++) l[(size_t)p * K * wl + e] = 0.0;
    for (int e = 0; e < K * wv; e++) v[(size_t)p * K * wv + e] = 0.0;
  }
  const double* Ap = A + (size_t)p * sA;
  const double* Bp = B + (size_t)p * sB;
  const double* cp = c + (size_t)p * sc;
  double* xp = x + (size_t)p * nx;
  // x_next = A x + B u + c  (A: nx x nx, B: nx x nu, column-major)
  double xn[32];
  for (int i = 0; i < nx; i++) {
    double s = cp ? cp[i] : 0.0;
    for (int j = 0; j < nx; j++) s += Ap[i + (size_t)j * nx] * xp[j];
    for (int j = 0; j < nu; j++) s += Bp[i + (size_t)j * nx] * up[j];
    xn[i] = s;
  }
  for (int j = 0; j < nu; j++) U[((size_t)p * steps + step) * nu + j] = up[j];
  for (int i = 0; i < nx; i++) {
    xp[i] = xn[i];
    X[((size_t)p * (steps + 1) + step + 1) * nx + i] = xn[i];
  }
}

__global__ void log_state_kernel(const double* x, double* X, int batch, int nx, int steps) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * nx) return;
  const int p = t / nx, i = t % nx;
  X[(size_t)p * (steps + 1) * nx + i] = x[t];
}

bool IsDev(const void* p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

struct fbstab_mpc_closed_loop {
  int N = 0, nx = 0, nu = 0, nc = 0, batch = 0, device = 0, shared = 0;
  fbstab_mpc_batch* solver = nullptr;
  double* data[12] = {};  // device copies of the 11 sequences and x0 (= current states)
  double *Asim = nullptr, *Bsim = nullptr, *csim = nullptr;  // explicit plant (else stage 0)
  double *x_init = nullptr, *z = nullptr, *l = nullptr, *v = nullptr, *y = nullptr;
  double* u_prev = nullptr;
  fbstab_out* out = nullptr;  // steps x batch
  double *X = nullptr, *U = nullptr;
  int steps_cap = 0, step = 0;
  std::vector<void*> owned;
  int Alloc(void** p, size_t bytes) {
    if (cudaMalloc(p, bytes ? bytes : 8) != cudaSuccess) {
      cudaGetLastError();
      return Fail(FBSTAB_ERR_ALLOC, "cudaMalloc failed in the closed-loop handle");
    }
    owned.push_back(*p);
    return FBSTAB_OK;
  }
};

extern "C" {

int fbstab_mpc_closed_loop_destroy(fbstab_mpc_closed_loop* h) {
  if (!h) return FBSTAB_OK;
  cudaSetDevice(h->device);
  for (void* p : h->owned) cudaFree(p);
  fbstab_mpc_batch_destroy(h->solver);
  delete h;
  return FBSTAB_OK;
}

int fbstab_mpc_closed_loop_create(int N, int nx, int nu, int nc, int batch, int device,
                                  int shared_data, const double* Q, const double* R,
                                  const double* S, const double* q, const double* r,
                                  const double* A, const double* B, const double* c,
                                  const double* E, const double* L, const double* d,
                                  const double* x_init, const double* Asim, const double* Bsim,
                                  int max_steps, fbstab_mpc_closed_loop** handle) {
  if (!handle) return Fail(FBSTAB_ERR_INVALID, "null handle pointer");
  *handle = nullptr;
  if (batch < 1 || max_steps < 1) return Fail(FBSTAB_ERR_INVALID, "batch and max_steps must be >= 1");
  if (nx > 32) return Fail(FBSTAB_ERR_INVALID, "closed loop: nx must be <= 32");
  auto* h = new fbstab_mpc_closed_loop;
  h->N = N, h->nx = nx, h->nu = nu, h->nc = nc, h->batch = batch, h->device = device;
  h->shared = shared_data ? 1 : 0;
  h->steps_cap = max_steps;
  int rc = fbstab_mpc_batch_create(N, nx, nu, nc, batch, device, &h->solver);
  if (rc) {
    delete h;
    return rc;
  }
  auto bail = [&](int code) {
    fbstab_mpc_closed_loop_destroy(h);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return bail(Fail(FBSTAB_ERR_CUDA, "cudaSetDevice failed"));
  const size_t K = N + 1, sx = nx, su = nu, sc_ = nc, Bn = batch;
  const size_t sizes[11] = {K * sx * sx, K * su * su, K * su * sx, K * sx,        K * su,     N * sx * sx,
                            N * sx * su, N * sx,      K * sc_ * sx, K * sc_ * su, K * sc_};
  const double* src[11] = {Q, R, S, q, r, A, B, c, E, L, d};
  const size_t copies = h->shared ? 1 : Bn;
  for (int k = 0; k < 11; k++) {
    if (!src[k]) return bail(Fail(FBSTAB_ERR_INVALID, "null input pointer"));
    const size_t bytes = copies * sizes[k] * sizeof(double);
    if ((rc = h->Alloc((void**)&h->data[k], bytes))) return bail(rc);
    if (cudaMemcpy(h->data[k], src[k], bytes, cudaMemcpyDefault) != cudaSuccess)
      return bail(Fail(FBSTAB_ERR_CUDA, "copy of the OCP data failed"));
  }
  if (!x_init) return bail(Fail(FBSTAB_ERR_INVALID, "null x_init"));
  const size_t nz = K * (sx + su), nl = K * sx, nv = K * sc_;
  if ((rc = h->Alloc((void**)&h->data[11], Bn * sx * 8)) || (rc = h->Alloc((void**)&h->x_init, Bn * sx * 8)) ||
      (rc = h->Alloc((void**)&h->z, Bn * nz * 8)) || (rc = h->Alloc((void**)&h->l, Bn * nl * 8)) ||
      (rc = h->Alloc((void**)&h->v, Bn * nv * 8)) || (rc = h->Alloc((void**)&h->y, Bn * nv * 8)) ||
      (rc = h->Alloc((void**)&h->u_prev, Bn * su * 8)) ||
      (rc = h->Alloc((void**)&h->out, (size_t)max_steps * Bn * sizeof(fbstab_out))) ||
      (rc = h->Alloc((void**)&h->X, Bn * (max_steps + 1) * sx * 8)) ||
      (rc = h->Alloc((void**)&h->U, Bn * (size_t)max_steps * su * 8)))
    return bail(rc);
  if (cudaMemcpy(h->x_init, x_init, Bn * sx * 8, cudaMemcpyDefault) != cudaSuccess)
    return bail(Fail(FBSTAB_ERR_CUDA, "copy of x_init failed"));
  if (Asim || Bsim) {
    if (!Asim || !Bsim) return bail(Fail(FBSTAB_ERR_INVALID, "Asim and Bsim go together"));
    if ((rc = h->Alloc((void**)&h->Asim, sx * sx * 8)) || (rc = h->Alloc((void**)&h->Bsim, sx * su * 8)))
      return bail(rc);
    if (cudaMemcpy(h->Asim, Asim, sx * sx * 8, cudaMemcpyDefault) != cudaSuccess ||
        cudaMemcpy(h->Bsim, Bsim, sx * su * 8, cudaMemcpyDefault) != cudaSuccess)
      return bail(Fail(FBSTAB_ERR_CUDA, "copy of the plant model failed"));
  }
  *handle = h;
  return FBSTAB_OK;
}

int fbstab_mpc_closed_loop_set_options(fbstab_mpc_closed_loop* h, const fbstab_options* o) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  return fbstab_mpc_batch_set_options(h->solver, o);
}

int fbstab_mpc_closed_loop_reset(fbstab_mpc_closed_loop* h, void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  CL_CUDA(cudaSetDevice(h->device));
  const size_t K = h->N + 1, Bn = h->batch;
  CL_CUDA(cudaMemcpyAsync(h->data[11], h->x_init, Bn * h->nx * 8, cudaMemcpyDeviceToDevice, s));
  CL_CUDA(cudaMemsetAsync(h->z, 0, Bn * K * (h->nx + h->nu) * 8, s));
  CL_CUDA(cudaMemsetAsync(h->l, 0, Bn * K * h->nx * 8, s));
  CL_CUDA(cudaMemsetAsync(h->v, 0, Bn * K * h->nc * 8, s));
  CL_CUDA(cudaMemsetAsync(h->u_prev, 0, Bn * h->nu * 8, s));
  h->step = 0;
  const int n = h->batch * h->nx;
  log_state_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->data[11], h->X, h->batch, h->nx, h->steps_cap);
  CL_CUDA(cudaGetLastError());
  return FBSTAB_OK;
}

int fbstab_mpc_closed_loop_step(fbstab_mpc_closed_loop* h, int warm_start, void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (h->step >= h->steps_cap) return Fail(FBSTAB_ERR_INVALID, "closed loop: max_steps reached");
  cudaStream_t s = (cudaStream_t)stream;
  CL_CUDA(cudaSetDevice(h->device));
  const int K = h->N + 1, Bn = h->batch, nx = h->nx, nu = h->nu, nc = h->nc;
  if (h->step == 0) {
    int rc = fbstab_mpc_closed_loop_reset(h, stream);
    if (rc) return rc;
  }
  if (!warm_start) {
    CL_CUDA(cudaMemsetAsync(h->z, 0, (size_t)Bn * K * (nx + nu) * 8, s));
    CL_CUDA(cudaMemsetAsync(h->l, 0, (size_t)Bn * K * nx * 8, s));
    CL_CUDA(cudaMemsetAsync(h->v, 0, (size_t)Bn * K * nc * 8, s));
  } else if (h->step > 0) {
    const long long n = (long long)Bn * (2 * nx + nu + nc);
    shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->z, h->l, h->v, Bn, K, nx + nu, nx, nc);
    CL_CUDA(cudaGetLastError());
  }
  fbstab_out* out = h->out + (size_t)h->step * Bn;
  double** D = h->data;
  int rc = h->shared
               ? fbstab_mpc_batch_solve_shared(h->solver, Bn, D[0], D[1], D[2], D[3], D[4], D[5], D[6],
                                               D[7], D[8], D[9], D[10], D[11], h->z, h->l, h->v, h->y,
                                               out, stream)
               : fbstab_mpc_batch_solve(h->solver, Bn, D[0], D[1], D[2], D[3], D[4], D[5], D[6], D[7],
                                        D[8], D[9], D[10], D[11], h->z, h->l, h->v, h->y, out,
                                        stream);
  if (rc) return rc;
  // plant: explicit (Asim, Bsim; no offset) or stage 0 of the plant's own OCP
  const size_t Nn = h->N;
  const double *A = h->Asim ? h->Asim : D[5], *B = h->Asim ? h->Bsim : D[6];
  const double* c = h->Asim ? nullptr : D[7];
  const bool per_plant = !h->Asim && !h->shared;
  plant_kernel<<<(Bn + 127) / 128, 128, 0, s>>>(
      A, B, c, per_plant ? Nn * nx * nx : 0, per_plant ? Nn * nx * nu : 0, per_plant ? Nn * nx : 0, Bn,
      nx, nu, K, D[11], h->u_prev, h->z, h->l, h->v, nx, nc, out, h->X, h->U, h->step,
      h->steps_cap);
  CL_CUDA(cudaGetLastError());
  h->step++;
  return FBSTAB_OK;
}

int fbstab_mpc_closed_loop_run(fbstab_mpc_closed_loop* h, int steps, int warm_start, double* X,
                               double* U, fbstab_out* out, void* stream) {
  if (!h) return Fail(FBSTAB_ERR_INVALID, "null handle");
  if (steps < 1 || steps > h->steps_cap) return Fail(FBSTAB_ERR_INVALID, "steps out of range");
  cudaStream_t s = (cudaStream_t)stream;
  h->step = 0;
  for (int t = 0; t < steps; t++) {
    int rc = fbstab_mpc_closed_loop_step(h, warm_start, stream);
    if (rc) return rc;
  }
  // logs: X is (batch, steps_cap + 1, nx) on the device; the caller gets (batch, steps + 1, nx)
  const size_t Bn = h->batch, nx = h->nx, nu = h->nu;
  bool host = false;
  if (X) {
    host = host || !IsDev(X);
    CL_CUDA(cudaMemcpy2DAsync(X, (steps + 1) * nx * 8, h->X, (h->steps_cap + 1) * nx * 8,
                              (steps + 1) * nx * 8, Bn, cudaMemcpyDefault, s));
  }
  if (U) {
    host = host || !IsDev(U);
    CL_CUDA(cudaMemcpy2DAsync(U, steps * nu * 8, h->U, (size_t)h->steps_cap * nu * 8, steps * nu * 8, Bn,
                              cudaMemcpyDefault, s));
  }
  if (out) {
    host = host || !IsDev(out);
    CL_CUDA(cudaMemcpyAsync(out, h->out, (size_t)steps * Bn * sizeof(fbstab_out), cudaMemcpyDefault, s));
  }
  if (host) CL_CUDA(cudaStreamSynchronize(s));
  return FBSTAB_OK;
}

const char* fbstab_mpc_closed_loop_path(const fbstab_mpc_closed_loop* h) {
  return h ? fbstab_mpc_batch_path(h->solver) : "";
}

}  // extern "C"

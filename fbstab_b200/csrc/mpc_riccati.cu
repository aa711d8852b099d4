// mpc_riccati.cu -- persistent CTA-per-instance kernel for MPC-structured QPs
// (BASELINE configs 3 and 4) and its host-side plan.
//
// One CTA owns one MPC instance for its whole solve and walks the horizon
// (FBstabMpc::Solve -> FBstabAlgorithm::Solve, fbstab_algorithm-impl.h:113-304,
// with RiccatiLinearSolver, riccati_linear_solver.cc:77-344, as the linear
// solver).  Everything the sequential Riccati chain touches lives in shared
// memory:
//   * the stage temporaries (Q~, R~, S~, inv(LL'), three small vectors);
//   * the stage's Q,R,S,A,B,E,L matrices, streamed global -> shared by the TMA
//     engine (cp.async.bulk + mbarrier) into a 3-slot ring two stages ahead of
//     the sweep -- the stage data never waits on an L2 round trip;
//   * the factor blocks L,M,AM,SM,P,SG: resident in shared memory when the
//     whole horizon fits, otherwise built in a ring slot, written behind the
//     sweep by a TMA bulk store and prefetched again by TMA for the forward /
//     backward substitution;
//   * the iterates, residual and step vectors when they fit.
// What does not fit goes to a per-CTA global workspace that stays L2 resident.
// The host picks the placement that keeps the most CTAs per SM in flight: the
// chain is latency bound, so instances in flight are the throughput lever.
//
// The kernel is persistent: CTAs pull instance indices from a global atomic
// counter, so converged / infeasible instances free their CTA at once (no host
// round trips, natural load balance over the iteration-count spread).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.cuh"
#include "engine_args.cuh"
#include "mpc_riccati.cuh"
#include "mpc_riccati.h"

namespace fbs {

namespace {

struct MpcArgs {
  MpcLayout lay;
  MpcData data;
  CommonArgs c;
  // *mismatch == 0 (device flag written by MpcSharedDetect before the launch):
  // every instance carries the stage data of instance 0, and all CTAs then read
  // that one copy -- L2 / L1 resident instead of a DRAM stream per instance
  const int* mismatch;
};

__host__ __device__ inline int Even(int n) { return (n + 1) & ~1; }

__device__ inline double* Take(double*& p, int n) {
  double* r = p;
  p += Even(n);
  return r;
}

__device__ inline void TakeVars(double*& p, int nz, int nl, int nv, Vars* v) {
  v->z = Take(p, nz);
  v->l = Take(p, nl);
  v->v = Take(p, nv);
  v->y = Take(p, nv);
}

#ifndef FBS_MPC_MINB128
#define FBS_MPC_MINB128 3
#endif
// MINB > 0: an instantiation compiled for more resident CTAs per SM (fewer
// registers per thread, some spills): chosen when it lets the whole batch run
// in fewer waves -- every instance of a batch takes about equally long, so a
// half-empty last wave costs a full instance time.
template <int KNX, int KNU, int KNC, int KT, int MINB = 0>
__global__ void __launch_bounds__(KT ? KT : 128,
                                  MINB ? MINB : (KT == 32 ? 16 : KT == 64 ? 8 : KT == 96 ? 4 : FBS_MPC_MINB128))
mpc_riccati_kernel(const __grid_constant__ MpcArgs a) {
  extern __shared__ __align__(16) double dyn_smem[];
  __shared__ double red[4 * kRedSlots];  // blockDim <= 128
  __shared__ __align__(8) unsigned long long bars[2 * kRing];
  __shared__ int s_inst;
  Team t{red};
  const MpcLayout& lay = a.lay;
  const CommonArgs& c = a.c;
  const int N = lay.N, nx = lay.nx, nu = lay.nu, nc = lay.nc;
  const int nz = lay.nz, nl = lay.nl, nv = lay.nv;
  const int K = N + 1;

  MpcProblem<KNX, KNU, KNC, KT> p;
  p.N = N;
  p.nx = nx;
  p.nu = nu;
  p.nc = nc;
  p.nz = nz;
  p.nl = nl;
  p.nv = nv;
  p.lay = lay;
  Buffers w;
  {
    // shared-memory carve; must mirror MpcPlanInit
    double* sp = dyn_smem;
    double* gp = c.ws + (size_t)blockIdx.x * c.ws_stride;
    const int m = nx > nu ? nx : nu;
    p.Qt = Take(sp, nx * nx);
    p.Rt = Take(sp, nu * nu);
    p.St = Take(sp, nu * nx);
    p.Linv = Take(sp, nx * nx);
    p.sa = Take(sp, m);
    p.sb = Take(sp, m);
    p.sc = Take(sp, m);
    p.dslot0 = sp;
    if (lay.data_ring) sp += (size_t)kDataRing * lay.SD;
    if (lay.fac_smem) {
      p.fac = sp;
      sp += (size_t)K * lay.FS;
      p.fslot0 = nullptr;
    } else {
      p.fslot0 = sp;
      sp += (size_t)kRing * lay.FS;
      p.fac = gp;
      gp += (size_t)K * lay.FS;
    }
    double*& v1 = lay.g1_smem ? sp : gp;
    p.gamma = Take(v1, nv);
    p.mus = Take(v1, nv);
    w.ri.z = Take(v1, nz);
    w.ri.l = Take(v1, nl);
    w.ri.v = Take(v1, nv);
    TakeVars(v1, nz, nl, nv, &w.dx);
    double*& v2 = lay.g2_smem ? sp : gp;
    TakeVars(v2, nz, nl, nv, &w.xk);
    TakeVars(v2, nz, nl, nv, &w.xi);
    TakeVars(v2, nz, nl, nv, &w.xp);
    p.Gam = w.dx.y;
    p.tv = w.dx.v;
  }
  p.bar0 = tma::smem_addr(&bars[0]);
  p.dphase = 0;
  p.fphase = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * kRing; s++) tma::mbar_init(tma::smem_addr(&bars[s]), 1);
    tma::fence_mbar_init();
  }
  __syncthreads();

  const bool shared_data = a.mismatch != nullptr && *a.mismatch == 0;
  for (;;) {
    if (threadIdx.x == 0) s_inst = atomicAdd(c.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= c.batch) break;
    const size_t i = shared_data ? 0 : (size_t)inst;
    p.Q = a.data.Q + i * K * nx * nx;
    p.R = a.data.R + i * K * nu * nu;
    p.S = a.data.S + i * K * nu * nx;
    p.q = a.data.q + i * K * nx;
    p.r = a.data.r + i * K * nu;
    p.A = a.data.A + i * N * nx * nx;
    p.B = a.data.B + i * N * nx * nu;
    p.c = a.data.c + i * N * nx;
    p.E = a.data.E + i * K * nc * nx;
    p.L = a.data.L + i * K * nc * nu;
    p.d = a.data.d + i * K * nc;
    p.x0 = a.data.x0 + (size_t)inst * nx;
    if (c.comp < 0) {
      const size_t io = (size_t)inst;
      solve_instance(t, p, c.opts, w, c.z + io * nz, c.l + io * nl, c.v + io * nv,
                     c.y + io * nv, c.out + inst);
    } else {
      RunComponent(t, p, c, inst, w);
    }
  }
}

typedef void (*MpcKernel)(const MpcArgs);
struct Variant {
  int nx, nu, nc, block;
  MpcKernel fn;
  int minb;  // > 0: the dense-occupancy instantiation of the shape
};
// Compile-time specialisations for the OCP shapes of the BASELINE configs
// (servo motor, double integrator, spacecraft, copolymerisation); every other
// shape runs the run-time-sized instantiation.
const Variant kVariants[] = {
    {4, 1, 4, 32, mpc_riccati_kernel<4, 1, 4, 32>},
    {4, 1, 4, 32, mpc_riccati_kernel<4, 1, 4, 32, 28>, 28},
    {2, 1, 6, 32, mpc_riccati_kernel<2, 1, 6, 32>},
    {2, 1, 6, 32, mpc_riccati_kernel<2, 1, 6, 32, 28>, 28},
    {6, 3, 12, 32, mpc_riccati_kernel<6, 3, 12, 32>},
    {6, 3, 12, 32, mpc_riccati_kernel<6, 3, 12, 32, 28>, 28},
    {6, 3, 12, 64, mpc_riccati_kernel<6, 3, 12, 64>},
    // 96 threads: 168 registers x 96 and 51 KB let FOUR instances share an SM
    // (128 threads: three; the recursion's critical path is one warp either way)
    {18, 5, 10, 96, mpc_riccati_kernel<18, 5, 10, 96>},
    {18, 5, 10, 128, mpc_riccati_kernel<18, 5, 10, 128>},
    {18, 5, 10, 64, mpc_riccati_kernel<18, 5, 10, 64>},
};
const MpcKernel kGeneric = mpc_riccati_kernel<0, 0, 0, 0>;

int EnvInt(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// Fills the size-dependent part of the layout and, for a given placement, the
// shared / global footprints.  Mirrors the carve in the kernel.
void Footprint(MpcLayout* L) {
  const int nx = L->nx, nu = L->nu, nc = L->nc, K = L->N + 1;
  const int nxx = nx * nx, nuu = nu * nu, nux = nu * nx;
  const int m = std::max(nx, nu);
  L->oLf = 0;
  L->oM = nxx;
  L->oAM = 2 * nxx;
  L->oSM = 3 * nxx;
  L->oP = 3 * nxx + nux;
  L->oSG = 3 * nxx + 2 * nux;
  L->FS = Even(3 * nxx + 2 * nux + nuu);
  int o = 0;
  auto sub = [&](int n) {
    const int at = o;
    o += Even(n + 1);
    return at;
  };
  L->oQ = sub(nxx);
  L->oR = sub(nuu);
  L->oS = sub(nux);
  L->oA = sub(nxx);
  L->oB = sub(nux);
  L->oE = sub(nc * nx);
  L->oL = sub(nc * nu);
  L->SD = o;
  size_t smem = 2 * (size_t)Even(nxx) + Even(nuu) + Even(nux) + 3 * (size_t)Even(m);
  size_t ws = 0;
  if (L->data_ring) smem += (size_t)kDataRing * L->SD;
  if (L->fac_smem) {
    smem += (size_t)K * L->FS;
  } else {
    smem += (size_t)kRing * L->FS;
    ws += (size_t)K * L->FS;
  }
  const size_t vars = (size_t)Even(L->nz) + Even(L->nl) + 2 * (size_t)Even(L->nv);
  const size_t g1 = 3 * (size_t)Even(L->nv) + Even(L->nz) + Even(L->nl) + vars;
  const size_t g2 = 3 * vars;
  (L->g1_smem ? smem : ws) += g1;
  (L->g2_smem ? smem : ws) += g2;
  L->smem_doubles = (int)smem;
  L->ws_doubles = ws + 2;
}

}  // namespace

namespace {
int PlanInitImpl(MpcPlan* p, int N, int nx, int nu, int nc, int max_batch, int sm_count,
                 const char** err, const Variant* force_variant, int force_ring);
}

int MpcPlanInit(MpcPlan* p, int N, int nx, int nu, int nc, int max_batch,
                int sm_count, const char** err) {
  int rc = PlanInitImpl(p, N, nx, nu, nc, max_batch, sm_count, err, nullptr, -1);
  if (rc || !EnvInt("FBSTAB_MPC_DENSE_OCCUPANCY", 1) || EnvInt("FBSTAB_MPC_GENERIC", 0) ||
      EnvInt("FBSTAB_MPC_BLOCK", 0) || EnvInt("FBSTAB_MPC_PLACE", -1) >= 0)
    return rc;
  const int waves = (max_batch + p->grid_max - 1) / std::max(p->grid_max, 1);
  if (waves < 2) return rc;
  for (const Variant& v : kVariants) {
    if (v.minb <= 0 || v.nx != nx || v.nu != nu || v.nc != nc) continue;
    // fewer registers per thread and the stage data read in place (no ring in
    // shared memory): more CTAs per SM; taken only if that saves a wave
    MpcPlan q;
    const char* e2 = "";
    if (PlanInitImpl(&q, N, nx, nu, nc, max_batch, sm_count, &e2, &v, 0) != FBSTAB_OK) {
      MpcPlanFree(&q);
      continue;
    }
    if ((max_batch + q.grid_max - 1) / q.grid_max < waves) {
      MpcPlanFree(p);
      *p = q;
      return FBSTAB_OK;
    }
    MpcPlanFree(&q);
  }
  return rc;
}

namespace {
int PlanInitImpl(MpcPlan* p, int N, int nx, int nu, int nc, int max_batch, int sm_count,
                 const char** err, const Variant* force_variant, int force_ring) {
  MpcLayout base;
  memset(&base, 0, sizeof(base));
  base.N = N;
  base.nx = nx;
  base.nu = nu;
  base.nc = nc;
  base.nz = (N + 1) * (nx + nu);
  base.nl = (N + 1) * nx;
  base.nv = (N + 1) * nc;
  base.data_ring = force_ring >= 0 ? force_ring : EnvInt("FBSTAB_MPC_RING", 1);
  p->block = EnvInt("FBSTAB_MPC_BLOCK", (nx + nu) <= 12 ? 32 : 64);
  if (p->block < 32 || p->block > 128 || p->block % 32) {
    *err = "FBSTAB_MPC_BLOCK must be 32, 64, 96 or 128";
    return FBSTAB_ERR_INVALID;
  }
  MpcKernel fn = kGeneric;
  bool special = false;
  if (!EnvInt("FBSTAB_MPC_GENERIC", 0)) {
    const int want_block = EnvInt("FBSTAB_MPC_BLOCK", 0);
    for (const Variant& v : kVariants) {
      if (v.nx != nx || v.nu != nu || v.nc != nc || v.minb > 0) continue;
      // the first instantiation of a shape is its default; a later one is
      // taken only when its block size was asked for
      if (!special || v.block == want_block) {
        fn = v.fn;
        p->block = v.block;
      }
      special = true;
    }
  }
  if (force_variant) {
    fn = force_variant->fn;
    p->block = force_variant->block;
    special = true;
  }
  p->kernel = (const void*)fn;
  // placements, most resident first: A everything in shared memory; B factor
  // blocks streamed; C iterates in global too; D only the sweep's working set
  const int place[4][3] = {{1, 1, 1}, {0, 1, 1}, {0, 1, 0}, {0, 0, 0}};
  const char* pname[4] = {"factors+vectors in smem", "vectors in smem, factors via TMA ring",
                          "step/residual in smem, factors via TMA ring",
                          "sweep working set in smem, factors via TMA ring"};
  const size_t cta_max = 232448;
  const int force = EnvInt("FBSTAB_MPC_PLACE", -1);
  int best = -1, best_ctas = 0;
  MpcLayout lay[4];
  int ctas[4];
  for (int k = 0; k < 4; k++) {
    lay[k] = base;
    lay[k].fac_smem = place[k][0];
    lay[k].g1_smem = place[k][1];
    lay[k].g2_smem = place[k][2];
    Footprint(&lay[k]);
    const size_t bytes = (size_t)lay[k].smem_doubles * 8;
    ctas[k] = 0;
    if (bytes > cta_max) continue;
    if (cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bytes) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas[k], p->kernel, p->block, bytes) !=
            cudaSuccess) {
      cudaGetLastError();
      ctas[k] = 0;
    }
  }
  // The Riccati chain is latency bound (profiles/r1_mpc_*): instances in
  // flight are the throughput lever, so take the placement with the most CTAs
  // per SM and, among equals, the most shared-memory resident one.
  if (force >= 0 && force < 4 && ctas[force] > 0) {
    best = force;
    best_ctas = ctas[force];
  } else {
    for (int k = 0; k < 4; k++)
      if (ctas[k] > best_ctas) {
        best = k;
        best_ctas = ctas[k];
      }
  }
  if (best < 0) {
    *err = "MPC stage working set does not fit the shared memory of one SM";
    return FBSTAB_ERR_INVALID;
  }
  p->lay = lay[best];
  p->smem_bytes = (size_t)p->lay.smem_doubles * 8;
  if (cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)p->smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the MPC kernel";
    return FBSTAB_ERR_CUDA;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->kernel, p->block,
                                                    p->smem_bytes) != cudaSuccess ||
      occ < 1) {
    cudaGetLastError();
    *err = "the MPC kernel does not fit on an SM";
    return FBSTAB_ERR_CUDA;
  }
  p->ctas_per_sm = std::min(occ, EnvInt("FBSTAB_MPC_MAX_CTAS", best_ctas));
  size_t grid = (size_t)sm_count * p->ctas_per_sm;
  grid = std::min<size_t>(grid, (size_t)std::max(max_batch, 1));
  p->grid_max = (int)grid;
  if (cudaMalloc(&p->ws, p->lay.ws_doubles * grid * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    *err = "cudaMalloc of the MPC workspace failed";
    return FBSTAB_ERR_ALLOC;
  }
  snprintf(p->name, sizeof(p->name),
           "mpc-riccati-cta%s (%s, %s, %d thr, %d CTA/SM, %zu KB smem)",
           special ? "<nx,nu,nc>" : "", pname[best],
           p->lay.data_ring ? "stage data via TMA ring" : "stage data in place", p->block,
           p->ctas_per_sm, p->smem_bytes / 1024);
  return FBSTAB_OK;
}
}  // namespace

void MpcPlanFree(MpcPlan* p) {
  if (p->ws) cudaFree(p->ws);
  p->ws = nullptr;
}

int MpcLaunch(const MpcPlan& p, int batch, const MpcData& data, double* z,
              double* l, double* v, double* y, fbstab_out* out,
              const fbstab_options& opts, int comp, const fbstab_component_io* io,
              int* counter, const int* mismatch, cudaStream_t stream) {
  MpcArgs a;
  a.mismatch = mismatch;
  a.lay = p.lay;
  a.data = data;
  a.c.batch = batch;
  a.c.z = z;
  a.c.l = l;
  a.c.v = v;
  a.c.y = y;
  a.c.out = out;
  a.c.ws = p.ws;
  a.c.ws_stride = p.lay.ws_doubles;
  a.c.counter = counter;
  a.c.vec_in_smem = p.lay.g2_smem;
  a.c.opts = opts;
  a.c.comp = comp;
  if (io)
    a.c.io = *io;
  else
    memset(&a.c.io, 0, sizeof(a.c.io));
  const int grid = std::min(batch, p.grid_max);
  ((MpcKernel)p.kernel)<<<grid, p.block, p.smem_bytes, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace fbs

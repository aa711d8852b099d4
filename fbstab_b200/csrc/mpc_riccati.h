// mpc_riccati.h -- host interface of the CTA-per-instance MPC path
// (kernel in mpc_riccati.cu, device policy in mpc_riccati.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "fbstab_b200.h"

namespace fbs {

// Sizes and placement decided on the host and passed to the kernel by value.
//
// Factor block of stage i (FS doubles, contiguous, FS even):
//   [ L(i) nx*nx | M(i) nx*nx | AM(i) nx*nx | SM(i) nu*nx | P(i) nx*nu | SG(i) nu*nu ]
// (the reference keeps six separate MatrixSequences, riccati_linear_solver.cc:35-55;
// one contiguous block per stage makes a stage's factor ONE bulk copy).
// Stage-data slot of the TMA ring (SD doubles): the stage's Q,R,S,A,B,E,L
// matrices, each in a sub-slot with one double of slack so that the copy keeps
// the 16-byte phase of its global source.
struct MpcLayout {
  int N, nx, nu, nc, nz, nl, nv;
  int FS, oLf, oM, oAM, oSM, oP, oSG;
  int SD, oQ, oR, oS, oA, oB, oE, oL;
  int data_ring;  // stage data streamed through the shared-memory ring (TMA)
  int fac_smem;   // factor blocks resident in shared memory (else global + ring)
  int g1_smem;    // gamma, mus, residual, step vectors in shared memory
  int g2_smem;    // iterates xk, xi, xp in shared memory
  int smem_doubles;   // dynamic shared memory per CTA, in doubles
  size_t ws_doubles;  // global workspace per CTA, in doubles
};

struct MpcPlan {
  MpcLayout lay;
  int block = 32;
  int ctas_per_sm = 0;
  int grid_max = 0;
  size_t smem_bytes = 0;
  double* ws = nullptr;
  const void* kernel = nullptr;  // selected instantiation
  char name[200];
};

struct MpcData {
  const double *Q, *R, *S, *q, *r, *A, *B, *c, *E, *L, *d, *x0;
};

// Chooses the placement, sizes the grid and allocates the workspace.
// Returns 0 or an FBSTAB_ERR_* code (message in *err).
int MpcPlanInit(MpcPlan* p, int N, int nx, int nu, int nc, int max_batch,
                int sm_count, const char** err);
void MpcPlanFree(MpcPlan* p);
int MpcLaunch(const MpcPlan& p, int batch, const MpcData& data, double* z,
              double* l, double* v, double* y, fbstab_out* out,
              const fbstab_options& opts, int comp, const fbstab_component_io* io,
              int* counter, const int* mismatch, cudaStream_t stream);

}  // namespace fbs

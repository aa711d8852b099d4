// sparse_lane.h -- host interface of the lane-per-instance sparse QP path (sparse_lane.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "fbstab_b200.h"
#include "sparse_symbolic.h"

namespace fbs {

// The integer tables of sparse_symbolic.h in device memory (one allocation).
struct SparseDev {
  int nz, nl, nv, n, nnzH, nnzG, nnzA, nnzK, nnzL;
  const int *Hr_ptr, *Hr_col, *Hr_val;
  const int *Gp, *Gi, *Gr_ptr, *Gr_col, *Gr_val;
  const int *Ap, *Ai, *Ar_ptr, *Ar_col, *Ar_val;
  const int *iperm, *Kp, *Ki, *Kkind, *Kidx, *Krow;
  const int *Lp, *Li, *Sp, *Sc, *St;
};

// Workspace of one instance (a column of a warp's lane-interleaved block), in doubles.
size_t SparseLaneWsDoublesPerLane(const SparseDev& d);
// Instances per warp (32, 16 or 8) and instances resident at a time for a batch.
int SparseLaneLanes(int batch, int sms);
int SparseLaneSlots(int batch, int sms);
// One launch solves `batch` instances (instance-major value arrays); `ws` holds
// lane_slots * SparseLaneWsDoublesPerLane doubles.  Returns 0 on success.
int SparseLaneLaunch(const SparseDev& d, int batch, int lane_slots, const double* Hx, const double* f,
                     const double* Gx, const double* h, const double* Ax, const double* b,
                     double* z, double* l, double* v, double* y, fbstab_out* out,
                     const fbstab_options& opts, double* ws, int* counter, cudaStream_t stream);

// ---- warp-per-instance path (sparse_team.cu): L and the LDL' work vectors in shared memory
size_t SparseTeamSmemBytes(const SparseDev& d);
size_t SparseTeamWsDoubles(const SparseDev& d);  // global workspace of one CTA, in doubles
// Resident single-warp CTAs per SM; 0 when the factor does not fit shared memory.
int SparseTeamCtasPerSm(const SparseDev& d);
int SparseTeamLaunch(const SparseDev& d, int batch, int ctas, const double* Hx, const double* f,
                     const double* Gx, const double* h, const double* Ax, const double* b,
                     double* z, double* l, double* v, double* y, fbstab_out* out,
                     const fbstab_options& opts, double* ws, int* counter, cudaStream_t stream);

}  // namespace fbs

// dense_small.h -- host interface of the warp-per-problem dense path
// (kernel in dense_small.cu).
#pragma once
#include <cuda_runtime.h>

#include "fbstab_b200.h"

namespace fbs {

struct DenseSmallPlan {
  bool enabled = false;
  const char* name = "generic";
  int nz = 0, nl = 0, nv = 0;
  int grid = 0;
  int* counter = nullptr;
  size_t smem = 0;
};

// Instances in flight per CTA of the warp kernel.
int DenseSmallWarpsPerCta();
// Enables the plan when (nz,nl,nv) fits the warp kernel; returns 0.
int DenseSmallInit(DenseSmallPlan* p, int nz, int nl, int nv, int sm_count,
                   int* counter);
int DenseSmallLaunch(const DenseSmallPlan& p, int batch, const double* H,
                     const double* f, const double* G, const double* h,
                     const double* A, const double* b, double* z, double* l,
                     double* v, double* y, fbstab_out* out,
                     const fbstab_options& opts, int comp,
                     const fbstab_component_io* io, cudaStream_t stream);

}  // namespace fbs

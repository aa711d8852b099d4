// dense_small.cuh -- placeholder until the warp-per-problem path lands.
#pragma once
#include "common.cuh"
namespace fbs {
struct DenseSmallPlan {
  bool enabled = false;
  const char* name = "generic";
};
inline int DenseSmallInit(DenseSmallPlan*, int, int, int, int, int*) { return 0; }
inline int DenseSmallLaunch(const DenseSmallPlan&, int, const double*, const double*,
                            const double*, const double*, const double*, const double*,
                            double*, double*, double*, double*, fbstab_out*,
                            const fbstab_options&, cudaStream_t) { return 1; }
}  // namespace fbs

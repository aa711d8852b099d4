// mpc_lane.h -- host interface of the lane-per-instance MPC path (mpc_lane.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "fbstab_b200.h"
#include "mpc_riccati.h"

namespace fbs {

// True when (nx,nu,nc) has a compile-time instantiation of the lane kernel.
bool MpcLaneSupported(int nx, int nu, int nc);
// Smallest batch, as a multiple of 16 instances per SM, that the lane kernel
// solves faster than the CTA kernel (measured per shape).
double MpcLaneCrossover(int nx, int nu, int nc);
// Dynamic shared memory of one single-warp CTA (the stage-block ring) and the number of
// such CTAs an SM holds.
size_t MpcLaneSmemBytes(int nx, int nu, int nc);
int MpcLaneWarpsPerSm(int N, int nx, int nu, int nc);
// Warps launched for `batch` instances on a device with `sms` SMs (and warps per CTA).
int MpcLaneWarps(int N, int nx, int nu, int nc, int batch, int sms, int* per_cta);
// Lane-interleaved workspace of one warp (32 instances), in doubles.
size_t MpcLaneWsDoublesPerWarp(int N, int nx, int nu, int nc);
// Launches min(max_warps, ceil(batch/32)) single-warp CTAs; `ws` holds
// max_warps * MpcLaneWsDoublesPerWarp doubles.  Returns 0 on success.
// `mismatch` (one int) and `sdata` (MpcLaneSharedDoubles doubles), both device
// memory or both nullptr: enable the common-stage-data fast path -- a pass over
// the inputs checks on the device whether every instance carries the stage
// data of instance 0, and the kernel then reads it from one shared array.
size_t MpcLaneSharedDoubles(int N, int nx, int nu, int nc);
// One pass over the inputs: *mismatch (device int) = 0 iff every instance of the
// batch carries the stage data (everything but x0) of instance 0.
int MpcSharedDetect(int N, int nx, int nu, int nc, int batch, const MpcData& data,
                    int* mismatch, cudaStream_t stream);
int MpcLaneLaunch(int N, int nx, int nu, int nc, int batch, int max_warps,
                  const MpcData& data, double* z, double* l, double* v, double* y,
                  fbstab_out* out, const fbstab_options& opts, double* ws, int* counter,
                  int* mismatch, double* sdata, cudaStream_t stream,
                  bool shared_known = false);

}  // namespace fbs

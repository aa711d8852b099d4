// microbench.cu -- measured FP64 peaks of the device the engine runs on.
//
// MEASURED_PEAKS.json (driver-written) carries HBM bandwidth and bf16 tensor
// throughput only; this path computes in FP64, so the roofline denominator is
// measured here: a register-resident DFMA loop and an FP64 mma.sync (DMMA)
// loop over all SMs, timed with CUDA events.
#include <cuda_runtime.h>

#include "fbstab_b200.h"

namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4,
         x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; k++) c[k][0] = c[k][1] = threadIdx.x + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) dmma884(c[k][0], c[k][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Even warps run the DFMA loop, odd warps the DMMA loop: shows whether the
// FP64 vector pipe and the FP64 tensor path overlap or share one datapath.
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters_fma, int iters_mma,
                                                    double a, double b) {
  double s = 0;
  if ((threadIdx.x >> 5) & 1) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k][0] = c[k][1] = threadIdx.x + k;
    for (int i = 0; i < iters_mma; i++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int k = 0; k < 8; k++) dmma884(c[k][0], c[k][1], a, b);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
  } else {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters_fma; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, b);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int fbstab_fp64_peak_concurrent(int device, double* total_tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return FBSTAB_ERR_NOGPU;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FBSTAB_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess)
    return FBSTAB_ERR_ALLOC;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // per-warp work sized so that both halves take about equally long when the
  // two paths run at their stand-alone peaks (64*64 DFMA flops vs 32*512 DMMA
  // flops per outer iteration at ~equal TFLOP/s)
  const int iters_fma = 4096, iters_mma = 1024;
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    mixed_kernel<<<blocks, threads>>>(out, iters_fma, iters_mma, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double warps = (double)blocks * (threads / 32) / 2.0;
  const double flops = 2.0 * 64.0 * iters_fma * warps * 32.0 +
                       2.0 * 256.0 * 32.0 * iters_mma * warps;
  *total_tflops = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? FBSTAB_OK : FBSTAB_ERR_CUDA;
}

extern "C" int fbstab_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return FBSTAB_ERR_NOGPU;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FBSTAB_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess)
    return FBSTAB_ERR_ALLOC;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_fma = 1e30f, best_mma = 1e30f;
  const int iters_fma = 4096, iters_mma = 2048;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters_fma, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_fma) best_fma = ms;
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters_mma, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_mma) best_mma = ms;
  }
  const double fma_flops = 2.0 * 64.0 * iters_fma * (double)blocks * threads;
  // one m8n8k4 = 8*8*4 MACs per warp
  const double mma_flops = 2.0 * 256.0 * 32.0 * iters_mma * (double)blocks * (threads / 32);
  *dfma_tflops = fma_flops / (best_fma * 1e-3) / 1e12;
  *dmma_tflops = mma_flops / (best_mma * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? FBSTAB_OK : FBSTAB_ERR_CUDA;
}

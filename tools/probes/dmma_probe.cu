// dmma_probe.cu -- how fast can mma.sync.m8n8k4.f64 go as a function of the
// resident warps per SM, with the operand pattern of dense_large's tile loop
// (16 independent accumulators, 4+4 fragments per k-step)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_probe tools/probes/dmma_probe.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp8(double* dst, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16(double* dst, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
// MODE 6: as 4 with 16-byte cp.async (4 per thread per chunk); 7: as 4 with TMA
// bulk row copies (128 B per row, 8 rows per lane 0..31 of warp 0... spread: lane 0 of
// every warp issues its share) tracked by an mbarrier per ring slot
// MODE 0: register operands only; 1: fragments re-read from shared memory every
// k-step; 2: as 1 plus the Gamma scaling (4 DMUL per k-step); 3: as 2 plus one
// __syncthreads per 4 k-steps (a chunk); 4: as 3 plus the cp.async staging
// traffic of a chunk (8 x 8-byte copies per thread from an L2-resident buffer
// into a second shared region, wait_group 1); 5: as 4 with generic (non-LDS) loads
template <int MODE>
__global__ void probe(double* out, int iters, double a0, double b0, const double* gsrc,
                      const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r8 = lane >> 2, c4 = lane & 3;
  for (int i = threadIdx.x; i < 192 * 20 + 16; i += blockDim.x) sm[i] = a0 + 1e-9 * i;
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = lane + a + b;
  double af[4], bf[4];
#pragma unroll
  for (int a = 0; a < 4; a++) { af[a] = a0 + a; bf[a] = b0 + a; }
  const int wm = (warp >> 1) & 3, wn = warp & 1;
  double* ring = sm + 4096;
  // 1024-byte aligned ring for the swizzled TMA boxes
  double* tring = (double*)((((size_t)(sm + 4096)) + 1023) & ~(size_t)1023);
  const double* smr = sm;
  if (MODE == 5) {  // defeat the address-space inference: generic loads
    asm volatile("" : "+l"(smr));
  }
  __shared__ unsigned long long bars[3];
  const int nwarps = blockDim.x >> 5;
  if (MODE == 7 || MODE == 9) {
    if (threadIdx.x == 0) {
      for (int b = 0; b < 3; b++)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[b])), "r"(MODE == 9 ? 1 : nwarps));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  for (int it = 0; it < iters; it++) {
    if (MODE >= 4 && MODE != 7) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    if (MODE == 9 && it >= 2) mbar_wait((unsigned)__cvta_generic_to_shared(&bars[(it - 2) % 3]), ((it - 2) / 3) & 1);
    if (MODE == 7 && it >= 2) mbar_wait((unsigned)__cvta_generic_to_shared(&bars[(it - 2) % 3]), ((it - 2) / 3) & 1);
    if (MODE >= 3) __syncthreads();
    if (MODE == 9) {
      // three 64-row x 16-k boxes (128 B rows, 128B swizzle) per chunk, one thread
      if (threadIdx.x == 0) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[it % 3]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3 * 64 * 128) : "memory");
        double* dst = tring + (it % 3) * 3072;
#pragma unroll
        for (int bx = 0; bx < 3; bx++) {
          const int c0 = 16 * (it & 63), c1 = 64 * bx + 256 * (it & 3);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
              ::"r"((unsigned)__cvta_generic_to_shared(dst + bx * 1024)), "l"(&tmap), "r"(c0), "r"(c1), "r"(bar)
              : "memory");
        }
      }
    }
    if (MODE == 7) {
      // 256 rows of 128 B per chunk, spread over the warps' lane 0..7
      double* dst = ring + (it % 3) * 5136;
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[it % 3]);
      const int rows_per_warp = 256 / nwarps;
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows_per_warp * 128) : "memory");
      __syncwarp();
      for (int r = lane; r < rows_per_warp; r += 32) {
        const int row = warp * rows_per_warp + r;
        bulk_g2s((unsigned)__cvta_generic_to_shared(dst + row * 20), gsrc + (size_t)(row + 128 * (it & 3)) * 1024 + 16 * (it & 63), 128, bar);
      }
    }
    double2 stg[8];
    if (MODE == 8) {
      const int p = lane & 7, rr = threadIdx.x >> 3;
      int q = 0;
      for (int row = rr; row < 256; row += blockDim.x >> 3, q++)
        stg[q] = *reinterpret_cast<const double2*>(gsrc + (size_t)(row + 128 * (it & 3)) * 1024 + 16 * (it & 63) + 2 * p);
    }
    if (MODE == 6) {
      double* dst = ring + (it % 3) * 5136;
      const int p = lane & 7, rr = threadIdx.x >> 3;  // 8 pairs per row
      for (int row = rr; row < 256; row += blockDim.x >> 3)
        cp16(dst + row * 20 + 2 * p, gsrc + (size_t)(row + 128 * (it & 3)) * 1024 + 16 * (it & 63) + 2 * p);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (MODE == 4 || MODE == 5) {
      double* dst = ring + (it % 3) * 5136;
      const int sidx = 8 * warp + r8;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        cp8(dst + sidx * 20 + 4 * q + c4, gsrc + (size_t)(sidx + 128 * (it & 7)) * 1024 + 16 * (it & 63) + 4 * q + c4);
        cp8(dst + 2560 + sidx * 20 + 4 * q + c4, gsrc + (size_t)(sidx + 128 * ((it + 3) & 7)) * 1024 + 16 * (it & 63) + 4 * q + c4);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      if (MODE >= 1) {
        const double g = MODE == 2 ? smr[192 * 20 + 4 * kk + c4] : 1.0;
#pragma unroll
        for (int a = 0; a < 4; a++) af[a] = smr[(32 * wm + 8 * a + r8) * 20 + 4 * kk + c4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const double v = smr[(128 + 32 * wn + 8 * b + r8) * 20 + 4 * kk + c4];
          bf[b] = MODE == 2 ? g * v : v;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    if (MODE == 8) {
      double* dst = ring + ((it + 1) % 3) * 5136;
      const int p = lane & 7, rr = threadIdx.x >> 3;
      int q = 0;
      for (int row = rr; row < 256; row += blockDim.x >> 3, q++)
        *reinterpret_cast<double2*>(dst + row * 20 + 2 * p) = stg[q];
    }
  }
  double s = 0;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) s += acc[a][b][0] + acc[a][b][1];
  if (MODE == 8) s += sm[4096 + lane];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(int sms, double* out, const double* gsrc, const CUtensorMap& tmap) {
  const int iters = 2000;
  const size_t smem = (4096 + 3 * 5136 + 256) * sizeof(double);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int warps : {8, 16}) {
    const int threads = warps > 32 ? warps * 16 : warps * 32;  // <= 1024 threads per block
    const int blocks_per_sm = warps > 32 ? 2 : 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      probe<MODE><<<sms * blocks_per_sm, threads, smem>>>(out, iters, 1.0000001, 1e-9, gsrc, tmap);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 256.0 * 64.0 * iters * (double)sms * warps;
    const double cyc_per_dmma_per_smsp = best * 1e-3 * 1.965e9 / (64.0 * iters * warps / 4.0);
    printf("mode %d  warps/SM %2d  %7.2f TFLOP/s  %6.2f cycles per DMMA per SMSP (at 1965 MHz)  [%s]\n",
           MODE, warps, flops / (best * 1e-3) / 1e12, cyc_per_dmma_per_smsp, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 2 * 1024);
  double* gsrc;
  cudaMalloc(&gsrc, sizeof(double) * 1024 * 1024);
  cudaMemset(gsrc, 0, sizeof(double) * 1024 * 1024);
  CUtensorMap tmap;
  {
    cuInit(0);
    cuuint64_t gdim[2] = {1024, 1024};           // k (contiguous), rows
    cuuint64_t gstr[1] = {1024 * sizeof(double)};
    cuuint32_t box[2] = {16, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, gsrc, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
  }
  run<0>(prop.multiProcessorCount, out, gsrc, tmap);
  run<1>(prop.multiProcessorCount, out, gsrc, tmap);
  run<2>(prop.multiProcessorCount, out, gsrc, tmap);
  run<3>(prop.multiProcessorCount, out, gsrc, tmap);
  run<4>(prop.multiProcessorCount, out, gsrc, tmap);
  run<5>(prop.multiProcessorCount, out, gsrc, tmap);
  run<6>(prop.multiProcessorCount, out, gsrc, tmap);
  run<7>(prop.multiProcessorCount, out, gsrc, tmap);
  run<8>(prop.multiProcessorCount, out, gsrc, tmap);
  run<9>(prop.multiProcessorCount, out, gsrc, tmap);
  return 0;
}

// tcgen05 tensor-core peak micro-benchmark for sm_100a: back-to-back tcgen05.mma (cta_group::1, M=128, N=256, K=32 bytes per
// instruction) on every SM, operands in shared memory (A optionally in TMEM), accumulator in TMEM.  Prints the measured
// dense throughput per kind so that the 8-bit roofline denominator is MEASURED instead of "2 x cuBLAS bf16".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_peak tools/microbench/mma_peak.cu
//   ./tools/microbench/mma_peak            (on the GPU box)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sageattention_b200/csrc/ptx.cuh"
using namespace sab;

// kind: 0 = i8 (s8 x s8 -> s32), 1 = f8f6f4 (e4m3 x e4m3 -> f32), 2 = f16 (bf16 x bf16 -> f32), 3 = f8f6f4 with A from TMEM
template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  constexpr int M = 128, N = 256;
  // A: 128 rows x 128 B (swizzle-128B K-major tile = 4 K-steps of 32 B), B: 256 rows x 128 B
  uint8_t* sA = smem;
  uint8_t* sB = smem + M * 128;
  for (int i = threadIdx.x; i < (M + N) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&holder);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = KIND == 0 ? make_idesc(2, 1, 1, M, N) : (KIND == 2 ? make_idesc(1, 1, 1, M, N) : make_idesc(1, 0, 0, M, N));
    const uint64_t dA = make_smem_desc<128>(smem_u32(sA));
    const uint64_t dB = make_smem_desc<128>(smem_u32(sB));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (KIND == 0) umma_i8_ss(tmem, dA + 2 * k, dB + 2 * k, idesc, 1);
          else if constexpr (KIND == 1) umma_f8_ss(tmem, dA + 2 * k, dB + 2 * k, idesc, 1);
          else if constexpr (KIND == 3) umma_f8_ts(tmem, tmem + 256 + 8 * k, dB + 2 * k, idesc, 1);
          else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(dA + 2 * k), "l"(dB + 2 * k), "r"(idesc), "r"(1) : "memory");
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int KIND>
static void run(const char* name, int kelems, int nsm) {
  long long* d;
  cudaMalloc(&d, nsm * sizeof(long long));
  const size_t smem = (128 + 256) * 128 + 1024;
  cudaFuncSetAttribute(mma_peak_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 20000;
  mma_peak_kernel<KIND><<<nsm, 128, smem>>>(200, d);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    mma_peak_kernel<KIND><<<nsm, 128, smem>>>(iters, d);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  long long h[256];
  cudaMemcpy(h, d, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  const double flops = 2.0 * 128 * 256 * kelems * 4.0 * iters * nsm;
  printf("%-34s %8.1f TFLOP/s dense  (%d SMs, %.3f ms, %.1f cycles per 128x256xK%d MMA, err=%s)\n", name, flops / (best * 1e-3) / 1e12,
         nsm, best, double(h[0]) / (4.0 * iters), kelems, cudaGetErrorString(err));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, nsm);
  run<0>("tcgen05.mma kind::i8 (SS)", 32, nsm);
  run<1>("tcgen05.mma kind::f8f6f4 e4m3 (SS)", 32, nsm);
  run<3>("tcgen05.mma kind::f8f6f4 e4m3 (TS)", 32, nsm);
  run<2>("tcgen05.mma kind::f16 bf16 (SS)", 16, nsm);
  return 0;
}

// MUFU.EX2 throughput on sm_100a, event-timed and clock64-timed: W warps per SM sub-partition, 16 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mufu_rate tools/microbench/mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 1) mufu_kernel(int iters, float seed, float* out, long long* cyc) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + 0.001f * float(i) + 1e-6f * float(threadIdx.x);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      x[i] = y * 0.25f - 1.5f;   // keeps the argument in (-1.5, 0.5): one FFMA between dependent MUFUs
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  float* dout; long long* dcyc;
  cudaMalloc(&dout, nsm * 1024 * 4); cudaMalloc(&dcyc, nsm * 8);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, nominal clock %d kHz\n", p.name, nsm, clk);
  for (int w : {1, 2, 4, 8}) {
    const int threads = 128 * w, iters = 20000;
    mufu_kernel<<<nsm, threads>>>(100, 0.1f, dout, dcyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mufu_kernel<<<nsm, threads>>>(iters, 0.1f, dout, dcyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost);
    const double mufu_per_smsp = double(iters) * 16 * w;           // warp-level MUFU instructions per sub-partition
    printf("%d warps/SMSP: %.3f ms, clock64 %lld cycles -> %.2f cycles per warp-MUFU per SMSP (clock64), %.2f (events @ %.3f GHz implied by clock64/ms); ex2 lanes/clk/SM = %.1f\n",
           w, ms, h[0], double(h[0]) / mufu_per_smsp, ms * 1e-3 * (double(h[0]) / (ms * 1e-3)) / mufu_per_smsp, double(h[0]) / (ms * 1e6),
           32.0 * 4 * mufu_per_smsp / double(h[0]));
  }
  return 0;
}

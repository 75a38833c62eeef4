// TMEM -> register bandwidth of tcgen05.ld by shape and by the number of warps loading at once (sm_100a).
// Each participating warp repeatedly loads the same 32 lanes x 64 columns x 4 B = 8 KB of ITS lane quarter and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/ldtm_bw tools/microbench/ldtm_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sageattention_b200/csrc/ptx.cuh"
using namespace sab;

#define R8(a, i) "=r"(a[i]), "=r"(a[i+1]), "=r"(a[i+2]), "=r"(a[i+3]), "=r"(a[i+4]), "=r"(a[i+5]), "=r"(a[i+6]), "=r"(a[i+7])
// SHAPE 0: 32x32b.x32 twice; 1: 32x32b.x64 once; 2: 16x256b.x8 (16 lanes x 64 columns) twice (lanes 0-15, 16-31);
// 3: 16x128b.x16 twice; 4: 16x64b.x32 twice; 5: 32x32b.x16 four times
template <int SHAPE>
__device__ __forceinline__ void load_8k(uint32_t t, uint32_t (&r)[64]) {
  if constexpr (SHAPE == 0) {
    tmem_ld32(t, *reinterpret_cast<uint32_t (*)[32]>(&r[0]));
    tmem_ld32(t + 32, *reinterpret_cast<uint32_t (*)[32]>(&r[32]));
  } else if constexpr (SHAPE == 1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24), R8(r, 32), R8(r, 40), R8(r, 48), R8(r, 56) : "r"(t) : "memory");
  } else if constexpr (SHAPE == 2) {
    // 16x256b.x8: 16 lanes x (8 x 256 bit = 64 columns); 32 registers per thread
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(t) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 32), R8(r, 40), R8(r, 48), R8(r, 56) : "r"(t + (16u << 16)) : "memory");
  } else if constexpr (SHAPE == 3) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(t) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 32), R8(r, 40), R8(r, 48), R8(r, 56) : "r"(t + (16u << 16)) : "memory");
  } else if constexpr (SHAPE == 4) {
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(t) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 32), R8(r, 40), R8(r, 48), R8(r, 56) : "r"(t + (16u << 16)) : "memory");
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) tmem_ld16(t + 16 * k, *reinterpret_cast<uint32_t (*)[16]>(&r[16 * k]));
  }
}

template <int SHAPE>
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, int nwarps, uint32_t* out, long long* cyc) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp < nwarps) {
    const uint32_t t = tmem + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[64], acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      load_8k<SHAPE>(t, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 64; i += 16) acc ^= r[i];
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int SHAPE>
static void run(const char* name, int nwarps, int nsm, uint32_t* dout, long long* dcyc) {
  const int iters = 2000;
  ldtm_kernel<SHAPE><<<nsm, 512>>>(20, nwarps, dout, dcyc);
  cudaDeviceSynchronize();
  ldtm_kernel<SHAPE><<<nsm, 512>>>(iters, nwarps, dout, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, dcyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double c = double(h[0]) / iters;
  printf("%-28s %2d warps loading: %7.1f cycles per 8 KB load+wait per warp -> %6.1f B/clk per warp, %6.1f B/clk per SM (%s)\n", name, nwarps, c,
         8192.0 / c, 8192.0 * nwarps / c, cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  uint32_t* dout; long long* dcyc;
  cudaMalloc(&dout, nsm * 512 * 4); cudaMalloc(&dcyc, nsm * 16 * 8);
  printf("%s, %d SMs; one CTA of 16 warps per SM\n", p.name, nsm);
  for (int w : {1, 2, 4, 8, 16}) {
    run<0>("32x32b.x32 x2", w, nsm, dout, dcyc);
    run<1>("32x32b.x64", w, nsm, dout, dcyc);
    run<5>("32x32b.x16 x4", w, nsm, dout, dcyc);
    run<2>("16x256b.x8 x2", w, nsm, dout, dcyc);
    run<3>("16x128b.x16 x2", w, nsm, dout, dcyc);
    run<4>("16x64b.x32 x2", w, nsm, dout, dcyc);
  }
  return 0;
}

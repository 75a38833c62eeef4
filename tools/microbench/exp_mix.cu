// Softmax inner-loop instruction-mix micro-benchmark for sm_100a: how close to the MUFU (ex2) roofline can W warps per SM
// sub-partition get when each of their 64-element "tiles" costs the attention kernel's per-element instruction mix
//   I2FP (int32 -> f32), FFMA2 (dequant, packed), MUFU.EX2, FADD2 (row sum, packed), F2FP e4m3x2 pack, VIMNMX3 (running int max)
// with everything in registers (no TMEM, no barriers)?  Prints cycles per tile and warp and the MUFU utilisation (8 cycles per
// warp-instruction at 4 lanes/clk per sub-partition) for several mixes, so that the kernel's losses can be split into
// "instruction mix" and "synchronisation structure".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/exp_mix tools/microbench/exp_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sageattention_b200/csrc/ptx.cuh"
using namespace sab;

// MIX bits: 1 = I2FP conversions, 2 = int max (VIMNMX3), 4 = e4m3 packing, 8 = row sum, 16 = int->float via IMAD magic (FMA pipe) instead of I2FP,
// 32 = a compiler scheduling barrier (empty volatile asm with a memory clobber) after every 16 elements, like the kernel's tcgen05.st / prefetch points
template <int MIX>
__global__ void __launch_bounds__(512, 1) exp_mix_kernel(int tiles, const int* __restrict__ in, uint32_t* __restrict__ out, long long* cyc, int one) {
  uint32_t s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = uint32_t(in[(threadIdx.x * 7 + i * 13) & 1023]);
  const float coef = 1.0e-4f + 1.0e-9f * float(threadIdx.x);
  float m = 3.0f, d = 0.f;
  int pm0 = -1000000000, pm1 = pm0, pm2 = pm0, pm3 = pm0;
  uint32_t keep = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    const uint64_t coef2 = pack_f2(coef, coef), nm2 = pack_f2(-m, -m);
    uint64_t acc0 = 0ull, acc1 = 0ull;
    uint32_t pk[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      float e[4];
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        const int i = 4 * w + u;
        float f0, f1;
        if constexpr ((MIX & 16) != 0) {   // bits of (1.5*2^23 + S) as a float: one IMAD-class add on the FMA pipe
          f0 = __uint_as_float(uint32_t(int(s[i]) * one) + 0x4B400000u);       // `one` == 1 at run time: forces IMAD (FMA pipe)
          f1 = __uint_as_float(uint32_t(int(s[i + 1]) * one) + 0x4B400000u);
        } else if constexpr ((MIX & 1) != 0) {
          f0 = __int2float_rn(int(s[i]));
          f1 = __int2float_rn(int(s[i + 1]));
        } else {
          f0 = __uint_as_float(s[i]);
          f1 = __uint_as_float(s[i + 1]);
        }
        float y0, y1;
        unpack_f2(ffma2(pack_f2(f0, f1), coef2, nm2), y0, y1);
        e[u] = ex2_approx(y0);
        e[u + 1] = ex2_approx(y1);
        if constexpr ((MIX & 8) != 0) {
          if (w & 1) acc1 = fadd2(acc1, pack_f2(e[u], e[u + 1]));
          else acc0 = fadd2(acc0, pack_f2(e[u], e[u + 1]));
        } else {
          keep ^= __float_as_uint(e[u]) ^ __float_as_uint(e[u + 1]);
        }
        if constexpr ((MIX & 2) != 0) {
          if (u == 0) { if (w & 1) pm0 = __vimax3_s32(pm0, int(s[i]), int(s[i + 1])); else pm1 = __vimax3_s32(pm1, int(s[i]), int(s[i + 1])); }
          else { if (w & 1) pm2 = __vimax3_s32(pm2, int(s[i]), int(s[i + 1])); else pm3 = __vimax3_s32(pm3, int(s[i]), int(s[i + 1])); }
        }
      }
      if constexpr ((MIX & 4) != 0) pk[w] = pack_e4m3x4(e[0], e[1], e[2], e[3]);
      if constexpr ((MIX & 32) != 0) {
        if ((w & 3) == 3) asm volatile("" : "+r"(pk[w]) : : "memory");
      }
    }
    if constexpr ((MIX & 4) != 0) {
#pragma unroll
      for (int w = 0; w < 16; ++w) keep ^= pk[w];
    }
    if constexpr ((MIX & 8) != 0) {
      float a0, a1, a2, a3;
      unpack_f2(acc0, a0, a1);
      unpack_f2(acc1, a2, a3);
      d += (a0 + a1) + (a2 + a3);
    }
    // next "tile": perturb the inputs so nothing is loop-invariant
#pragma unroll
    for (int i = 0; i < 64; i += 8) s[i] += keep & 1u;
    m += 1.0e-7f;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = keep ^ __float_as_uint(d) ^ uint32_t(pm0 ^ pm1 ^ pm2 ^ pm3);
}

template <int MIX>
static void run(const char* name, int warps_per_smsp, int nsm, const int* din, uint32_t* dout, long long* dcyc) {
  const int threads = 128 * warps_per_smsp, tiles = 2000;
  exp_mix_kernel<MIX><<<nsm, threads>>>(50, din, dout, dcyc, 1);
  cudaDeviceSynchronize();
  exp_mix_kernel<MIX><<<nsm, threads>>>(tiles, din, dout, dcyc, 1);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, dcyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  const double per_tile = double(h[0]) / tiles;                       // cycles for `warps_per_smsp` tiles per sub-partition
  const double mufu = 64.0 * 8.0 * warps_per_smsp / per_tile;         // fraction of the MUFU peak
  printf("%-46s %d warps/SMSP: %7.1f cycles per round of tiles, %6.1f per warp-tile, MUFU %5.1f %% (%s)\n", name, warps_per_smsp, per_tile,
         per_tile / warps_per_smsp, 100.0 * mufu, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  int* din; uint32_t* dout; long long* dcyc;
  cudaMalloc(&din, 1024 * 4); cudaMalloc(&dout, nsm * 1024 * 4); cudaMalloc(&dcyc, nsm * 8);
  int hin[1024];
  for (int i = 0; i < 1024; ++i) hin[i] = (i * 2654435761u) % 200000 - 100000;
  cudaMemcpy(din, hin, sizeof(hin), cudaMemcpyHostToDevice);
  printf("%s, %d SMs; one CTA per SM\n", p.name, nsm);
  for (int w : {1, 2, 4}) {
    run<0>("ex2 only (+FFMA2)", w, nsm, din, dout, dcyc);
    run<1>("I2FP + ex2", w, nsm, din, dout, dcyc);
    run<1 | 8>("I2FP + ex2 + row sum", w, nsm, din, dout, dcyc);
    run<1 | 4 | 8>("I2FP + ex2 + row sum + e4m3 pack", w, nsm, din, dout, dcyc);
    run<1 | 2 | 4 | 8>("full mix (I2FP, ex2, sum, pack, int max)", w, nsm, din, dout, dcyc);
    run<16 | 2 | 4 | 8>("full mix, IMAD-magic instead of I2FP", w, nsm, din, dout, dcyc);
    run<16 | 4 | 8>("IMAD-magic, no int max", w, nsm, din, dout, dcyc);
    run<1 | 2 | 4 | 8 | 32>("full mix + scheduling barrier every 16 elements", w, nsm, din, dout, dcyc);
  }
  return 0;
}

// Does concurrent tcgen05.mma activity slow the softmax instruction mix (MUFU.EX2 via the MIO path)?  One CTA per SM:
// warps 0..W-1 run the attention kernel's per-element mix on registers (see exp_mix.cu), the LAST warp optionally issues
// back-to-back tcgen05.mma (kind::i8 SS 128x64xK128 + kind::f8f6f4 TS 128x128xK64 per "tile", the attention kernel's QK + PV).
// Optional: tcgen05.ld / tcgen05.st traffic inside the exp loop like the kernel (S loads, P stores).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/exp_mma tools/microbench/exp_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sageattention_b200/csrc/ptx.cuh"
using namespace sab;

// MODE bits: 1 = MMA warp active, 2 = tcgen05.ld of 64 columns per tile + tcgen05.st x4 every 16 elements in the exp warps,
// 4 = (with 2) the load of the NEXT tile is issued inside the loop into a second register set (software prefetch), 8 = (with 2 and 4) in x16 pieces,
// 16 = all exponentials by the FMA-pipe polynomial (ptx.cuh ex2_poly2) instead of MUFU.EX2, 32 = no tcgen05.st (P kept in registers)
template <int MODE>
__global__ void __launch_bounds__(288, 1) exp_mma_kernel(int tiles, int nexp_warps, const int* __restrict__ in, uint32_t* __restrict__ out, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 8) tmem_alloc<512>(&holder);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 8) {
    if constexpr ((MODE & 1) != 0) {
      constexpr uint32_t idesc_qk = make_idesc(2, 1, 1, 128, 64), idesc_pv = make_idesc(1, 0, 0, 128, 128);
      const uint64_t dQ = make_smem_desc<128>(smem_u32(smem)), dK = make_smem_desc<128>(smem_u32(smem + 16384)),
                     dV = make_smem_desc<64>(smem_u32(smem + 32768));
      uint32_t ph = 0;
      while (!stop) {
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {     // 8 "tiles" of QK + PV between two commits
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_i8_ss(tmem + (r & 1) * 64, dQ + 2 * k, dK + 2 * k, idesc_qk, k > 0);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_f8_ts(tmem + 128, tmem + 320 + 8 * k, dV + 2 * k, idesc_pv, 1);
          }
          tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
    }
  } else if (warp < nexp_warps) {
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    uint32_t s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = uint32_t(in[(threadIdx.x * 7 + i * 13) & 1023]);
    const float coef = 1.0e-4f + 1.0e-9f * float(threadIdx.x);
    float m = 3.0f, d = 0.f;
    int pm0 = -1000000000, pm1 = pm0, pm2 = pm0, pm3 = pm0;
    uint32_t keep = 0;
    const long long t0 = clock64();
    uint32_t nx[64];
    if constexpr ((MODE & 4) != 0) {
      uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&nx[0]);
      uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&nx[32]);
      tmem_ld32(tmem + lane_off + 384, lo);
      tmem_ld32(tmem + lane_off + 416, hi);
    }
    for (int t = 0; t < tiles; ++t) {
      if constexpr ((MODE & 4) != 0) {
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; ++i) { asm volatile("" : "+r"(nx[i])); s[i] = (nx[i] & 0xffffu) + uint32_t(t); }
      } else if constexpr ((MODE & 2) != 0) {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld32(tmem + lane_off + 384, lo);
        tmem_ld32(tmem + lane_off + 416, hi);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; ++i) s[i] = (s[i] & 0xffffu) + uint32_t(t);
      }
      const uint64_t coef2 = pack_f2(coef, coef), nm2 = pack_f2(-m, -m);
      uint64_t acc0 = 0ull, acc1 = 0ull;
      uint32_t pk4[4];
#pragma unroll
      for (int w = 0; w < 16; ++w) {
        if constexpr ((MODE & 4) != 0 && (MODE & 8) == 0) {
          if (w == 4) {
            uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&nx[0]);
            uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&nx[32]);
            tmem_ld32(tmem + lane_off + 384, lo);
            tmem_ld32(tmem + lane_off + 416, hi);
          }
        }
        if constexpr ((MODE & 8) != 0) {
          if ((w & 3) == 0) {
            uint32_t (&pc)[16] = *reinterpret_cast<uint32_t (*)[16]>(&nx[4 * w]);
            tmem_ld16(tmem + lane_off + 384 + 4 * w, pc);
          }
        }
        float e[4];
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
          const int i = 4 * w + u;
          const float f0 = __int2float_rn(int(s[i])), f1 = __int2float_rn(int(s[i + 1]));
          float y0, y1;
          unpack_f2(ffma2(pack_f2(f0, f1), coef2, nm2), y0, y1);
          if constexpr ((MODE & 16) != 0) {
            ex2_poly2(y0, y1, e[u], e[u + 1]);
          } else {
            e[u] = ex2_approx(y0);
            e[u + 1] = ex2_approx(y1);
          }
          if (w & 1) acc1 = fadd2(acc1, pack_f2(e[u], e[u + 1]));
          else acc0 = fadd2(acc0, pack_f2(e[u], e[u + 1]));
          if (u == 0) { if (w & 1) pm0 = __vimax3_s32(pm0, int(s[i]), int(s[i + 1])); else pm1 = __vimax3_s32(pm1, int(s[i]), int(s[i + 1])); }
          else { if (w & 1) pm2 = __vimax3_s32(pm2, int(s[i]), int(s[i + 1])); else pm3 = __vimax3_s32(pm3, int(s[i]), int(s[i + 1])); }
        }
        pk4[w & 3] = pack_e4m3x4(e[0], e[1], e[2], e[3]);
        if ((w & 3) == 3) {
          if constexpr ((MODE & 2) != 0 && (MODE & 32) == 0) tmem_st4(tmem + lane_off + 448 + (w >> 2) * 4, pk4[0], pk4[1], pk4[2], pk4[3]);
          else keep ^= pk4[0] ^ pk4[1] ^ pk4[2] ^ pk4[3];
        }
      }
      float a0, a1, a2, a3;
      unpack_f2(acc0, a0, a1);
      unpack_f2(acc1, a2, a3);
      d += (a0 + a1) + (a2 + a3);
      if constexpr ((MODE & 2) != 0 && (MODE & 32) == 0) tc_wait_st();
      else {
#pragma unroll
        for (int i = 0; i < 64; i += 8) s[i] += keep & 1u;
      }
      m += 1.0e-7f;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = keep ^ __float_as_uint(d) ^ uint32_t(pm0 ^ pm1 ^ pm2 ^ pm3);
    if (warp == 0) { __syncwarp(); if (threadIdx.x == 0) { __threadfence_block(); stop = 1; } }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

template <int MODE>
static void run(const char* name, int nexp, int nsm, const int* din, uint32_t* dout, long long* dcyc) {
  const int tiles = 2000;
  cudaFuncSetAttribute(exp_mma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  exp_mma_kernel<MODE><<<nsm, 288, 64 * 1024>>>(50, nexp, din, dout, dcyc);
  cudaDeviceSynchronize();
  exp_mma_kernel<MODE><<<nsm, 288, 64 * 1024>>>(tiles, nexp, din, dout, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, dcyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%-58s %d exp warps: %7.1f cycles per tile and warp (%s)\n", name, nexp, double(h[0]) / tiles, cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  int* din; uint32_t* dout; long long* dcyc;
  cudaMalloc(&din, 1024 * 4); cudaMalloc(&dout, nsm * 1024 * 4); cudaMalloc(&dcyc, nsm * 8);
  int hin[1024];
  for (int i = 0; i < 1024; ++i) hin[i] = (i * 2654435761u) % 200000 - 100000;
  cudaMemcpy(din, hin, sizeof(hin), cudaMemcpyHostToDevice);
  printf("%s, %d SMs; one CTA per SM; full softmax mix per 64-element tile\n", p.name, nsm);
  for (int w : {4, 8}) {
    run<0>("registers only, tensor core idle", w, nsm, din, dout, dcyc);
    run<1>("registers only, MMA warp issuing QK+PV back to back", w, nsm, din, dout, dcyc);
    run<2>("+ tcgen05.ld S / tcgen05.st P per tile, tensor core idle", w, nsm, din, dout, dcyc);
    run<3>("+ tcgen05.ld / st, MMA warp issuing QK+PV back to back", w, nsm, din, dout, dcyc);
    run<6>("tcgen05.ld of the NEXT tile issued inside the loop (prefetch)", w, nsm, din, dout, dcyc);
    run<14>("prefetch in four x16 pieces spread over the loop", w, nsm, din, dout, dcyc);
    run<7>("prefetch + MMA warp", w, nsm, din, dout, dcyc);
    run<16>("registers only, polynomial exp2 on the FMA pipe (no MUFU)", w, nsm, din, dout, dcyc);
    run<22>("prefetch, polynomial exp2 (no MUFU)", w, nsm, din, dout, dcyc);
    run<54>("prefetch, polynomial exp2, no tcgen05.st", w, nsm, din, dout, dcyc);
    run<38>("prefetch, MUFU, no tcgen05.st", w, nsm, din, dout, dcyc);
  }
  return 0;
}

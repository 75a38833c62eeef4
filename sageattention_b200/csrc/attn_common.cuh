// Shared declarations of the attention kernels (attn_alt.cu: head_dim 128 product kernel; attn_hd64.cu: head_dim 64;
// attn.cu: exact-max / fp16-PV / masked / sequence-parallel variants and the host dispatch).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"
#include "ptx.cuh"

#ifdef SAB_TIMELINE
#define SAB_TL(slot) do { if (tl_on && j >= 16 && j < 48) tl[(j - 16) * 16 + (slot)] = clock64(); } while (0)
#else
#define SAB_TL(slot) do {} while (0)
#endif

namespace sab {

constexpr int BM = 128;  // Q rows per CTA
constexpr int BN = 64;   // keys per softmax / MMA tile (reference CTA_K)
constexpr int LK = 128;  // keys per TMA stage (two tiles)
constexpr uint32_t kTmemCols = 256;
constexpr float kFp8Offset = 8.807f;      // attn_utils.cuh:30
constexpr float kMaskValue = -5000000.0f; // attn_utils.cuh:310
constexpr int kIntSentinel = -(1 << 30);
#ifdef SAB_NO_TMA_STORE   // A/B switch: direct per-row global stores in the epilogue
constexpr bool kTmaStoreEpilogue = false;
#else
constexpr bool kTmaStoreEpilogue = true;
#endif
constexpr int kAlphaCol = 40;  // column of an S buffer (beyond the 16 / 32 P columns) that carries alpha(j) to the correction warps

struct AttnParams {
  const float* q_scale;
  const float* k_scale;
  const float* v_scale;  // nullable
  const float* v_mean;   // nullable
  void* out;
  float* lse;  // nullable
  int64_t o_stride_b, o_stride_h, o_stride_s;
  int B, Hq, Hkv, Sq, Sk;
  int n_q_tiles;
  int causal;
  float sm_scale_log2;
  int q_mult;            // scales per 128-row Q block: 1 / 4 / 32
  int q_gran;            // 1 / 2 / 3
  int k_mult;            // scales per 64-key block: 1 / 4
  int64_t qs_stride_bh;  // dense: scales per (b,h); varlen: unused
  int64_t ks_stride_bh;
  int qs_stride_idx;     // dense 1, varlen Hq
  int ks_stride_idx;     // dense 1, varlen Hkv
  int ks_vec4;           // per-thread K scales, dense, 16-byte aligned: the four scales of a key tile load as one float4
  const int32_t* cu_q;   // varlen (nullable)
  const int32_t* cu_k;
  const int32_t* cu_v;
  const int32_t* cu_qs;
  const int32_t* cu_ks;
  int causal_q_offset;   // global index of query row 0 (sequence-parallel causal)
  int kv_seg_len;        // > 0: K/V are rank-major all-gathered segments of this many keys
  int32_t* dbg;          // nullable debug dump (CTA 0 only)
  // attention mask of the Triton-path shell (attn_qk_int8_per_block.py:33-52): [B,Hq,Sq,Skv] with arbitrary (also 0) strides
  const void* mask;      // nullable
  int mask_kind;         // 1: bool (uint8, false = masked out), 2: additive bias in the output dtype (fp16 / bf16)
  int64_t mask_sb, mask_sh, mask_sm, mask_sn;   // element strides
  // sequence-parallel gather fused into the launch (kSeg instantiations): K/V segment `seg` of KV-head group g is complete in
  // this GPU's memory once seg_flags[g * n_segments + seg] == seg_epoch (written in stream order behind the peer copies)
  const uint32_t* seg_flags;   // nullable
  uint32_t seg_epoch;
  int seg_heads;               // KV heads per flag group
  // dense outputs: byte-typed 4-D map of O (box 128 B x 128 rows, 128-byte swizzle) for the TMA-store epilogue of attn_alt.cu
  int o_tma;                   // 1: o_map is valid
  int n_items;                 // attn_q4.cu: length of the work list (Q tiles x heads x batch)
  CUtensorMap o_map;
};

__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
#ifdef SAB_WATCHDOG
  // Debug build: bounded spin.  A wait that never completes reports itself (block 0 only) and is ABANDONED, so the kernel
  // runs to the end with wrong results instead of hanging the GPU, and the printf buffer is flushed.
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0)
        printf("sab: mbarrier timeout warp %d bar@%u parity %u\n", threadIdx.x >> 5, smem_u32(bar) & 0xfffu, parity);
      break;
    }
  }
#else
  mbar_wait(bar, parity);
#endif
}

// Host: opt a kernel into its dynamic shared-memory size once PER DEVICE (cudaFuncSetAttribute applies to the current
// device only; a process that drives several GPUs — sageattn_host(device=...), tests — must set it on each).
template <typename Kern>
inline int ensure_dynamic_smem(Kern kern, size_t smem, bool (&done)[64]) {
  int dev = 0;
  SAB_CUDA_OK(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < 64;
  if (!tracked || !done[dev]) {
    SAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    if (tracked) done[dev] = true;
  }
  return SAB_OK;
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}


}  // namespace sab

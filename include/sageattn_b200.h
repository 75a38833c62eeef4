/*
 * sageattn_b200 — C ABI of the B200-native (sm_100a) SageAttention hot path.
 *
 * Plain C: raw device pointers, sizes, element strides and a cudaStream_t (passed as void*).
 * No torch types.  Every entry point returns 0 on success or a negative sab_status; the message of
 * the last failure on the calling thread is available from sab_last_error().  Nothing is allocated
 * or retained by the library: the caller owns every buffer (the reference convention).
 *
 * Each function names the reference interface it replaces (paths relative to thu-ml/SageAttention):
 *   csrc/fused/fused.h:19-76            the eight `_fused` quantisation entry points
 *   csrc/qattn/attn_cuda_sm89.h:19-104  the seven `_qattn_sm89` attention entry points
 *   sageattention/triton/quant_per_thread.py:154-203, quant_per_block.py:49-101,
 *   quant_per_block_varlen.py:60-104    the Triton quantisation launchers
 *   sageattention/triton/attn_qk_int8_block_varlen.py:123-150 (+ causal)   the varlen attention launcher
 *
 * Tensor convention: a "bhsd" tensor is addressed as ptr[b*stride_b + h*stride_h + s*stride_s + d]
 * (strides in ELEMENTS, innermost dimension contiguous) which covers both reference layouts
 * ("HND" [B,H,S,D] and "NHD" [B,S,H,D], tensor_layout 1 / 0 in the csrc/qattn launchers).
 */
#ifndef SAGEATTN_B200_H_
#define SAGEATTN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SAB_OK = 0,
  SAB_ERR_INVALID = -1,     /* bad argument (reference: TORCH_CHECK / std::invalid_argument) */
  SAB_ERR_UNSUPPORTED = -2, /* unsupported head_dim / dtype / granularity                     */
  SAB_ERR_CUDA = -3,        /* CUDA runtime / driver failure (launch, tensor-map encode)       */
  SAB_ERR_ARCH = -4         /* device is not sm_100 (no fallback path exists)                  */
} sab_status;

/* element types of the 16-bit inputs / outputs */
#define SAB_DTYPE_FP16 0
#define SAB_DTYPE_BF16 1

/* QuantGranularity, csrc/qattn/attn_utils.cuh:52-58 (ints 2/3 cross the reference op boundary) */
#define SAB_GRAN_PER_BLOCK 1
#define SAB_GRAN_PER_WARP 2
#define SAB_GRAN_PER_THREAD 3

/* rounding / epsilon semantics of the INT8 quantisers */
#define SAB_SEM_CUDA 0   /* csrc/fused/fused.cu:147-184: amax floored 1e-7, x*(127/amax), cvt.rni       */
#define SAB_SEM_TRITON 1 /* triton/quant_per_block.py:41-45: scale=amax/127, x/scale, round-half-away    */

const char* sab_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else SAB_ERR_ARCH. */
int sab_check_device(void);
int sab_version(void);

/* ------------------------------------------------------------------------------------------------
 * K smoothing mean.  Replaces `km = k.mean(dim=seq, keepdim=True)` (sageattention/core.py:773,
 * :433 for varlen): fp32 accumulation, result rounded to the input dtype.  mean is [B,H,D]
 * contiguous in `dtype`.  For the varlen form pass B=1, S=total tokens (mean over all sequences).
 * workspace: >= sab_k_mean_workspace_bytes(B,H,S,D) bytes of device scratch.
 * ---------------------------------------------------------------------------------------------- */
int64_t sab_k_mean_workspace_bytes(int B, int H, int S, int D);
int sab_k_mean(const void* k, int dtype, void* mean, int B, int H, int S, int D, int64_t stride_b,
               int64_t stride_h, int64_t stride_s, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-block INT8 quantisation of one tensor.  Replaces
 *   quant_per_block_int8_cuda (both overloads), quant_per_block_int8_fuse_sub_mean_cuda,
 *   quant_per_warp_int8_cuda            (csrc/fused/fused.h:19-49; semantics SAB_SEM_CUDA)
 *   triton quant_per_block_int8_kernel  (triton/quant_per_block.py:21-47; SAB_SEM_TRITON)
 * Every `blk` consecutive rows of a (b,h) share one scale; scale is [B,H,scale_cols] fp32 with
 * scale_cols >= ceil(S/blk) (the reference pads per_warp scales to ceil(S/BLKQ)*(BLKQ/WARPQ); the
 * padding entries are written with the all-zero-rows value).
 *   mean     : optional [B,H,D] (`dtype`), subtracted before quantisation.  SAB_SEM_CUDA subtracts
 *              in fp32 (fused.cu:119-126); SAB_SEM_TRITON first rounds (x-mean) to `dtype`
 *              (`k = k - km` in torch, quant_per_block.py:53-54).
 *   sm_scale : multiplied in fp32 before quantisation when has_sm_scale != 0.
 * blk in {16,32,64,128}; D in {64,128}.
 * ---------------------------------------------------------------------------------------------- */
int sab_quant_per_block_int8(const void* x, int dtype, const void* mean, int8_t* out, float* scale,
                             int B, int H, int S, int D, int64_t x_stride_b, int64_t x_stride_h,
                             int64_t x_stride_s, int64_t o_stride_b, int64_t o_stride_h,
                             int64_t o_stride_s, int scale_cols, int blk, int semantics,
                             int has_sm_scale, float sm_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-thread INT8 quantisation.  Replaces triton quant_query_per_thread_int8_kernel /
 * quant_key_per_thread_int8_kernel (triton/quant_per_thread.py:21-98).
 *   is_key == 0: inside each 32-row block, group g = rows {g, g+8, g+16, g+24}; 8 scales / block;
 *                scale [B,H,ceil(S/128)*4*8].
 *   is_key != 0: inside each 64-key block, group t = keys {8j+2t, 8j+2t+1}; 4 scales / block;
 *                scale [B,H,ceil(S/64)*4].
 * scale = amax/127 + 1e-7, round-half-away; `mean` (optional) handled as SAB_SEM_TRITON above.
 * ---------------------------------------------------------------------------------------------- */
int sab_quant_per_thread_int8(const void* x, int dtype, const void* mean, int8_t* out, float* scale,
                              int B, int H, int S, int D, int64_t x_stride_b, int64_t x_stride_h,
                              int64_t x_stride_s, int64_t o_stride_b, int64_t o_stride_h,
                              int64_t o_stride_s, int scale_cols, int is_key, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Varlen per-block INT8 quantisation.  Replaces triton/quant_per_block_varlen.py:21-58.
 * x, out: packed [T,H,D] (token stride = stride_t, head stride = stride_h); cu_seqlens [nseq+1]
 * int32 device; cu_scale [nseq+1] int32 device = exclusive cumsum of ceil(L_i/blk); scale is
 * [cu_scale[nseq], H] (block-major, head-minor).  Semantics SAB_SEM_TRITON.
 * ---------------------------------------------------------------------------------------------- */
int sab_quant_per_block_int8_varlen(const void* x, int dtype, const void* mean, int8_t* out, float* scale,
                                    const int32_t* cu_seqlens, const int32_t* cu_scale, int nseq,
                                    int max_seqlen, int H, int D, int64_t x_stride_t, int64_t x_stride_h,
                                    int64_t o_stride_t, int64_t o_stride_h, int blk, int has_sm_scale,
                                    float sm_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-channel FP8 (e4m3) quantisation of V.  Replaces transpose_pad_permute_cuda +
 * scale_fuse_quant_cuda / mean_scale_fuse_quant_cuda (csrc/fused/fused.h:58-76,
 * sageattention/quant.py:224-293).
 *   v      : bhsd, `dtype`.
 *   v_fp8  : [B,H,D,S_pad] bytes, token-contiguous ("V transposed"), S_pad = ceil(S/128)*128, tokens
 *            >= S are written as 0.  Unlike the reference there is NO 16-token permutation: that is
 *            an mma.sync fragment artefact (fused.cu:287-291); tcgen05 reads V^T tiles directly.
 *   v_scale: [B,H,D] fp32 = amax/scale_max;  v_fp8 = cvt.rn.satfinite.e4m3(v*(scale_max/amax)).
 *   v_mean : optional [B,H,D] fp32 (smooth_v): mean over tokens (denominator: S rounded up to 16,
 *            fused.cu:349,381), subtracted before quantisation.
 * Varlen form (cu_seqlens != NULL): v packed [T,H,D] (B ignored, stride_b ignored, stride_s = token
 * stride), per-channel statistics over ALL tokens, output [H,D,T_pad] where sequence i starts at
 * token column cu_pad[i] (cu_pad = exclusive cumsum of ceil(L_i/128)*128, int32 device, nseq+1).
 * workspace: >= sab_per_channel_fp8_workspace_bytes(...) bytes.
 * ---------------------------------------------------------------------------------------------- */
int64_t sab_per_channel_fp8_workspace_bytes(int B, int H, int S, int D);
int sab_per_channel_fp8(const void* v, int dtype, uint8_t* v_fp8, float* v_scale, float* v_mean, int B,
                        int H, int S, int D, int64_t stride_b, int64_t stride_h, int64_t stride_s,
                        int64_t s_pad, float scale_max, const int32_t* cu_seqlens,
                        const int32_t* cu_pad, int nseq, int max_seqlen, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Single-pass front-end ("fuse quantisation into producers").  One launch each; a thread-block
 * cluster per (b,h) reduces the statistic through distributed shared memory and then quantises, so K / V
 * are read from HBM once (the second read hits L2) instead of twice:
 *  sab_k_smooth_quant_int8   = sab_k_mean + sab_quant_per_thread_int8(is_key) [SAB_GRAN_PER_THREAD] or
 *                              sab_k_mean + sab_quant_per_block_int8(blk 64, SAB_SEM_CUDA) [SAB_GRAN_PER_WARP / PER_BLOCK],
 *                              i.e. `km = k.mean(...)` + the K half of per_thread_int8 / per_warp_int8
 *                              (sageattention/core.py:773, 788-795); mean_out [B,H,D] in `dtype` is also returned
 *                              (needed for the LSE correction, core.py:782-786).  Dense layouts only.
 *  sab_per_channel_fp8_fused = sab_per_channel_fp8 (dense form) without workspace.
 * The INT8 / FP8 outputs and scales are bit-identical to the two-step entry points given the same mean.
 * ---------------------------------------------------------------------------------------------- */
int sab_k_smooth_quant_int8(const void* k, int dtype, void* mean_out, int8_t* out, float* scale, int B, int H, int S,
                            int D, int64_t x_stride_b, int64_t x_stride_h, int64_t x_stride_s, int64_t o_stride_b,
                            int64_t o_stride_h, int64_t o_stride_s, int scale_cols, int granularity, void* stream);
int sab_per_channel_fp8_fused(const void* v, int dtype, uint8_t* v_fp8, float* v_scale, float* v_mean, int B, int H,
                              int S, int D, int64_t stride_b, int64_t stride_h, int64_t stride_s, int64_t s_pad,
                              float scale_max, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Building blocks of the sequence-parallel path (no reference counterpart: the reference ships no SP
 * code; semantics chosen so that the sharded result equals the single-GPU one):
 *  sab_channel_stats   : per-(b,h,d) fp32 sum / max / min over the LOCAL tokens (all-reduced by the host:
 *                        SUM for the global K mean, MAX/MIN for the global per-channel V range).
 *  sab_v_quant_with_amax: sab_per_channel_fp8 with the per-channel |max| supplied by the caller.
 * ---------------------------------------------------------------------------------------------- */
int sab_channel_stats(const void* x, int dtype, float* sum_out, float* max_out, float* min_out, int B, int H,
                      int S, int D, int64_t stride_b, int64_t stride_h, int64_t stride_s, void* workspace,
                      void* stream);
int sab_v_quant_with_amax(const void* v, int dtype, uint8_t* v_fp8, const float* amax, float* v_scale, int B,
                          int H, int S, int D, int64_t stride_b, int64_t stride_h, int64_t stride_s,
                          int64_t s_pad, float scale_max, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention: INT8 QK^T (tcgen05 kind::i8) -> fp32 online softmax -> FP8 P -> FP8 PV
 * (tcgen05 kind::f8f6f4, fp32 accumulation in tensor memory) -> fp16/bf16 O.
 * Replaces qk_int8_sv_f8_accum_{f32,f16}[_fuse_v_scale][_fuse_v_mean]_attn[_inst_buf]
 * (csrc/qattn/attn_cuda_sm89.h:19-104) and, with cu_seqlens, the Triton varlen launchers.
 *   q_int8 : bhsd int8 [B,Hq,Sq,D];  k_int8: bhsd int8 [B,Hkv,Skv,D];
 *   v_fp8  : [B,Hkv,D,s_pad] e4m3 as produced by sab_per_channel_fp8;
 *   out    : bhsd `out_dtype` [B,Hq,Sq,D];  lse: optional [B,Hq,Sq] fp32 (log2 units, as the
 *            reference kernel writes it, qk_int_sv_f8_cuda_sm89.cuh:691-703).
 *   q_scale/k_scale : fp32, packed exactly as the reference packs them for the given granularity
 *            (q: [B,Hq,ceil(Sq/128)*{1,4,32}], k: [B,Hkv,ceil(Skv/64)*{1,1,4}]);
 *   v_scale: optional [B,Hkv,D] (fuse_v_scale); v_mean: optional [B,Hkv,D] (fuse_v_mean).
 *   sm_scale: softmax scale; the kernel multiplies by log2(e) itself.  Pass fold_sm_scale != 0 when
 *            sm_scale*log2e is already folded into q (per_block Triton path, core.py:304).
 *   is_causal: top-left aligned (kv_idx > q_idx masked, attn_utils.cuh:310).
 * Sequence-parallel form (kv_seg_len > 0, dense only): k_int8 is [P*B,Hkv,kv_seg_len,D] and v_fp8 is
 * [P*B,Hkv,D,kv_seg_len] exactly as an all-gather of the P ranks' local shards lays them out (rank-major);
 * key t of batch b lives in segment t / kv_seg_len.  Skv = P*kv_seg_len, kv_seg_len % 128 == 0.
 * causal_q_offset: global index of this call's first query row (top-left causal alignment on global
 * indices); 0 on a single GPU.
 * Varlen form (cu_seqlens_q != NULL): q/out packed [Tq,Hq,D], k packed [Tk,Hkv,D] (B = number of
 * sequences, *_stride_b ignored), v_fp8 [Hkv,D,T_pad] with cu_pad_v offsets, scales
 * [nblocks_total,H] with cu_*_scale offsets (quant_per_block_varlen.py:72-79), per_block only.
 * ---------------------------------------------------------------------------------------------- */
int sab_qk_int8_sv_f8_attn(const int8_t* q_int8, const int8_t* k_int8, const uint8_t* v_fp8, void* out,
                           float* lse, const float* q_scale, const float* k_scale, const float* v_scale,
                           const float* v_mean, int out_dtype, int B, int Hq, int Hkv, int Sq, int Skv,
                           int D, int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s,
                           int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_s, int64_t v_s_pad,
                           int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s, int is_causal,
                           int q_gran, int k_gran, float sm_scale, int fold_sm_scale,
                           const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                           const int32_t* cu_pad_v, const int32_t* cu_q_scale, const int32_t* cu_k_scale,
                           int max_seqlen_q, int causal_q_offset, int kv_seg_len, int32_t* debug_dump,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * FP16-PV variant: INT8 QK^T -> fp32 online softmax WITHOUT exponent offset -> FP16 P -> FP16 PV
 * (tcgen05 kind::f16, fp32 accumulation in tensor memory).  Replaces the Triton attention launchers
 * `forward` of sageattention/triton/attn_qk_int8_per_block.py:130-183, attn_qk_int8_per_block_causal.py,
 * attn_qk_int8_block_varlen.py:123-150, attn_qk_int8_per_block_causal_varlen.py (the kernels behind
 * sageattn_qk_int8_pv_fp16_triton and sageattn_varlen).  Arguments as sab_qk_int8_sv_f8_attn except:
 *   v_f16t : [B,Hkv,D,s_pad] fp16 (token-contiguous, zero padded) from sab_v_transpose_f16;
 *   no v_scale / v_mean (V is not quantised); no sequence-parallel form.
 * sab_v_transpose_f16 replaces `v.to(torch.float16)` (sageattention/core.py:297-298) plus the K-major
 * re-layout tcgen05 needs; varlen form as sab_per_channel_fp8.
 * ---------------------------------------------------------------------------------------------- */
int sab_v_transpose_f16(const void* v, int dtype, void* v_f16t, int B, int H, int S, int D, int64_t stride_b,
                        int64_t stride_h, int64_t stride_s, int64_t s_pad, const int32_t* cu_seqlens,
                        const int32_t* cu_pad, int nseq, int max_seqlen, void* stream);
int sab_qk_int8_sv_f16_attn(const int8_t* q_int8, const int8_t* k_int8, const void* v_f16t, void* out, float* lse,
                            const float* q_scale, const float* k_scale, int out_dtype, int B, int Hq, int Hkv,
                            int Sq, int Skv, int D, int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s,
                            int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_s, int64_t v_s_pad,
                            int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s, int is_causal, int q_gran,
                            int k_gran, float sm_scale, int fold_sm_scale, const int32_t* cu_seqlens_q,
                            const int32_t* cu_seqlens_k, const int32_t* cu_pad_v, const int32_t* cu_q_scale,
                            const int32_t* cu_k_scale, int max_seqlen_q, void* stream);

/* Masked form of sab_qk_int8_sv_f16_attn: the `attn_mask` argument of sageattn_qk_int8_pv_fp16_triton
 * (sageattention/core.py:160-331: bool or q.dtype, broadcast to [B,Hq,Sq,Skv], non-causal only) as consumed by
 * sageattention/triton/attn_qk_int8_per_block.py:33-52:
 *   SAB_MASK_BOOL : uint8 elements, 0 = masked out (the reference adds -1e6 and skips all-false blocks);
 *   SAB_MASK_BIAS : additive bias in the output dtype (fp16 / bf16), added to S*scale in fp32 before the softmax.
 * mask_stride_* are ELEMENT strides of the broadcast view (0 for broadcast dimensions).  Dense, non-causal, K scales
 * per block / per warp; a row whose keys are all masked returns zeros (the reference: an unmasked softmax). */
#define SAB_MASK_BOOL 1
#define SAB_MASK_BIAS 2
int sab_qk_int8_sv_f16_attn_masked(const int8_t* q_int8, const int8_t* k_int8, const void* v_f16t, void* out, float* lse,
                                   const float* q_scale, const float* k_scale, int out_dtype, int B, int Hq, int Hkv,
                                   int Sq, int Skv, int D, int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s,
                                   int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_s, int64_t v_s_pad,
                                   int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s, int q_gran, int k_gran,
                                   float sm_scale, int fold_sm_scale, const void* attn_mask, int mask_kind,
                                   int64_t mask_stride_b, int64_t mask_stride_h, int64_t mask_stride_m,
                                   int64_t mask_stride_n, void* stream);

/* Sequence-parallel attention with the K/V gather fused into the launch (no reference counterpart: the reference ships no
 * sequence-parallel operator, example/parallel_sageattn_cogvideo.py:44-51 only wires sageattn into xfuser).  Same tensors as the
 * kv_seg_len > 0 form of sab_qk_int8_sv_f8_attn (k_int8 [P*B,Hkv,kv_seg_len,D], v_fp8 [P*B,Hkv,D,kv_seg_len], rank-major), but
 * the segments may still be arriving: the kernel's TMA producer waits until
 *     seg_flags[(kv_head / heads_per_flag) * (Skv / kv_seg_len) + segment] == seg_epoch      (32-bit, system-scope acquire)
 * before it loads the first tile of a segment.  The caller writes each flag in stream order behind the copies that fill that
 * (KV-head group, segment) — e.g. cudaMemcpyAsync from peer memory followed by cuStreamWriteValue32 on a second stream — and
 * passes a fresh epoch per call, so flags never need resetting.  The copies and flag writes must not need SMs that the
 * waiting CTAs occupy (copy engines + stream memory operations do not).  Dense, non-causal. */
int sab_qk_int8_sv_f8_attn_sp(const int8_t* q_int8, const int8_t* k_int8, const uint8_t* v_fp8, void* out, float* lse,
                              const float* q_scale, const float* k_scale, const float* v_scale, int out_dtype, int B, int Hq,
                              int Hkv, int Sq, int Skv, int D, int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s,
                              int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_s, int64_t v_s_pad, int64_t o_stride_b,
                              int64_t o_stride_h, int64_t o_stride_s, int q_gran, int k_gran, float sm_scale, int kv_seg_len,
                              const uint32_t* seg_flags, uint32_t seg_epoch, int heads_per_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAGEATTN_B200_H_ */

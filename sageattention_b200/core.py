"""Public API — drop-in for sageattention/core.py of thu-ml/SageAttention, backed by ONE sm_100a backend.

Signatures and behaviour follow the reference (file:line cited per function); the bodies re-state the
host logic (asserts, head-dim padding, K smoothing, LSE correction, output slicing) around the
B200 kernels: fused quantisers (csrc/quant.cu) and the tcgen05 attention kernel (csrc/attn.cu).
There is no arch dispatch, no Triton, no CPU fallback.
"""
from typing import Any, Optional
import warnings
import torch
import torch.nn.functional as F

from . import ops
from ._capi import SAB_GRAN_PER_BLOCK, SAB_GRAN_PER_WARP, SAB_GRAN_PER_THREAD
from .quant import (k_mean, per_block_int8, per_warp_int8, per_thread_int8, per_channel_fp8, quant_q_int8, quant_k_int8, smooth_quant_k,
                    per_block_int8_varlen, per_channel_fp8_varlen, transpose_v_f16, transpose_v_f16_varlen)

_LOG2E = 1.44269504


def _check_inputs(q, k, v):
    assert q.is_cuda, "Input tensors must be on cuda."
    assert q.dtype in [torch.float16, torch.bfloat16], "Input tensors must be in dtype of torch.float16 or torch.bfloat16"
    assert q.device == k.device == v.device, "All tensors must be on the same device."
    assert q.dtype == k.dtype == v.dtype, "All tensors must have the same dtype."


def _pad_head_dim(q, k, v):
    """sageattention/core.py:752-761."""
    head_dim_og = q.size(-1)
    if head_dim_og < 64:
        pad = 64 - head_dim_og
    elif 64 < head_dim_og < 128:
        pad = 128 - head_dim_og
    elif head_dim_og > 128:
        raise ValueError(f"Unsupported head_dim: {head_dim_og}")
    else:
        pad = 0
    if pad:
        q, k, v = F.pad(q, (0, pad)), F.pad(k, (0, pad)), F.pad(v, (0, pad))
    return q, k, v, head_dim_og


def _lse_correction(q, km, tensor_layout):
    """q @ km^T in the input dtype (sageattention/core.py:775-786)."""
    nh_dim = 2 if tensor_layout == "NHD" else 1
    g = q.size(nh_dim) // km.size(nh_dim)
    kmb = torch.repeat_interleave(km, g, dim=nh_dim) if g > 1 else km
    if tensor_layout == "NHD":
        return torch.matmul(q.transpose(1, 2), kmb.transpose(1, 2).transpose(2, 3)).squeeze(-1).to(torch.float32)
    return torch.matmul(q, kmb.transpose(2, 3)).squeeze(-1).to(torch.float32)


def sageattn_qk_int8_pv_fp8_cuda(
    q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, tensor_layout: str = "HND", is_causal: bool = False,
    qk_quant_gran: str = "per_thread", sm_scale: Optional[float] = None, pv_accum_dtype: str = "fp32+fp16",
    smooth_k: bool = True, smooth_v: bool = False, return_lse: bool = False, **kwargs: Any,
) -> torch.Tensor:
    """INT8 QK^T + FP8 PV attention (reference: sageattention/core.py:636-826).

    pv_accum_dtype keeps the reference vocabulary.  On B200 the PV product always accumulates in fp32
    inside the tensor core (tcgen05 f32 accumulation is full-rate, so the reference's f16 first-level
    accumulator — a consumer-GPU speed trick — buys nothing); the option still selects the reference's V
    quantisation range: 2.25 for "fp32+fp16", 448 otherwise (core.py:805-807).  smooth_v is honoured
    only for "fp32" (core.py:797-803).  Unknown values raise (the reference silently returns an
    uninitialised tensor)."""
    dtype = q.dtype
    _check_inputs(q, k, v)
    assert qk_quant_gran in ["per_warp", "per_thread"], "qk_quant_gran must be either 'per_warp' or 'per_thread'."
    if pv_accum_dtype not in ("fp32", "fp32+fp32", "fp32+fp16"):
        raise ValueError(f"Unsupported pv_accum_dtype: {pv_accum_dtype}")

    _tensor_layout = 0 if tensor_layout == "NHD" else 1
    if tensor_layout not in ("NHD", "HND"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    q, k, v, head_dim_og = _pad_head_dim(q, k, v)
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of qkv must be contiguous."
    if is_causal:
        seq_dim = 1 if _tensor_layout == 0 else 2
        assert q.size(seq_dim) == k.size(seq_dim), "qo_len and kv_len must be equal for causal attention."
    if sm_scale is None:
        sm_scale = head_dim_og ** -0.5

    # front-end: K mean (2 launches), Q / K quantisers, V statistics (2) + FP8 quantiser — 7 launches (reference: 6 + k.mean, core.py:773-809).
    # The single-launch cluster variants (quant.smooth_quant_k, per_channel_fp8(fused=True)) exist but measured slower on B200.
    lse_correction = None
    km = k_mean(k, tensor_layout) if smooth_k else None
    if smooth_k and return_lse:
        lse_correction = _lse_correction(q, km, tensor_layout)
    q_int8, q_scale = quant_q_int8(q, qk_quant_gran, tensor_layout)
    k_int8, k_scale = quant_k_int8(k, km, qk_quant_gran, tensor_layout)
    gran = SAB_GRAN_PER_WARP if qk_quant_gran == "per_warp" else SAB_GRAN_PER_THREAD

    o = torch.empty(q.size(), dtype=dtype, device=q.device)

    if pv_accum_dtype in ("fp32+fp32", "fp32+fp16") and smooth_v:
        warnings.warn(f"pv_accum_dtype is '{pv_accum_dtype}', smooth_v will be ignored.")
        smooth_v = False
    quant_v_scale_max = 2.25 if pv_accum_dtype == "fp32+fp16" else 448.0
    v_fp8, v_scale, vm = per_channel_fp8(v, tensor_layout=tensor_layout, scale_max=quant_v_scale_max, smooth_v=smooth_v)

    lse = ops.qk_int8_sv_f8_attn(q_int8, k_int8, v_fp8, o, q_scale, k_scale, v_scale, vm, _tensor_layout,
                                 1 if is_causal else 0, gran, gran, sm_scale, 0, 1 if return_lse else 0)
    o = o[..., :head_dim_og]
    if return_lse:
        return o, lse / _LOG2E + lse_correction * sm_scale if smooth_k else lse / _LOG2E
    return o


def sageattn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, tensor_layout: str = "HND", is_causal: bool = False,
             sm_scale: Optional[float] = None, return_lse: bool = False, **kwargs: Any):
    """sageattention/core.py:79-157.  The reference dispatches on compute capability and raises on sm_100;
    here the single B200 backend is used with the reference's sm89 defaults (per-thread INT8, FP8 PV,
    "fp32+fp16" V range).  Unknown kwargs (attn_mask=, dropout_p=, scale=, ...) are accepted and ignored
    exactly like the reference so SDPA monkey-patches keep working (example/cogvideox_infer.py:35).

    The 2.25 V range of "fp32+fp16" exists in the reference to keep its fp16 PV accumulator from overflowing; this kernel
    accumulates PV in fp32, so the range buys nothing here and is kept ONLY so that the quantised V bytes (and with them the
    output) match what the reference's default sm89 dispatch produces.  Callers who prefer e4m3's full dynamic range for V call
    sageattn_qk_int8_pv_fp8_cuda(..., pv_accum_dtype="fp32+fp32") (448 range, the reference's sm90 choice) at the same speed."""
    return sageattn_qk_int8_pv_fp8_cuda(q, k, v, tensor_layout=tensor_layout, is_causal=is_causal, sm_scale=sm_scale,
                                        return_lse=return_lse, pv_accum_dtype="fp32+fp16")


def sageattn_qk_int8_pv_fp8_cuda_sm90(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                                      qk_quant_gran: str = "per_thread", sm_scale: Optional[float] = None,
                                      pv_accum_dtype: str = "fp32+fp32", smooth_k: bool = True, smooth_v: bool = False,
                                      return_lse: bool = False, **kwargs: Any):
    """API shell for sageattention/core.py:829-996 (Hopper entry point): same numerics contract, served by
    the sm_100a kernel.  Only "fp32+fp32" exists in the reference for this entry (core.py:985-989)."""
    assert pv_accum_dtype == "fp32+fp32", "only 'fp32+fp32' is supported for the sm90 entry point"
    return sageattn_qk_int8_pv_fp8_cuda(q, k, v, tensor_layout=tensor_layout, is_causal=is_causal,
                                        qk_quant_gran=qk_quant_gran, sm_scale=sm_scale, pv_accum_dtype=pv_accum_dtype,
                                        smooth_k=smooth_k, smooth_v=False, return_lse=return_lse)


def sageattn_qk_int8_pv_fp16_triton(q, k, v, tensor_layout: str = "HND", quantization_backend: str = "triton",
                                    is_causal: bool = False, sm_scale: Optional[float] = None, smooth_k: bool = True,
                                    return_lse: bool = False, attn_mask: Optional[torch.Tensor] = None, **kwargs: Any):
    """sageattention/core.py:160-331 on sm_100a: per-block INT8 quantisation with the Triton path's exact rounding
    (bit-exact q/k/scales), sm_scale*log2e folded into q, then the FP16-PV kernel variant (tcgen05 kind::f16, P and
    V in fp16, softmax without exponent offset — the Triton kernel's numerics, triton/attn_qk_int8_per_block.py).
    attn_mask (bool, or the q dtype as an additive bias; broadcast to [B,Hq,Sq,Skv]; non-causal only) follows
    core.py:248-250, 310-325 and attn_qk_int8_per_block.py:33-52 — including the reference's convention that a float
    bias is added AFTER the logits were scaled by sm_scale*log2(e), i.e. it acts as bias*ln(2) in natural-log units."""
    dtype = q.dtype
    _check_inputs(q, k, v)
    if attn_mask is not None:
        assert attn_mask.dtype == torch.bool or attn_mask.dtype == q.dtype, "attn_mask must be of dtype bool or the same dtype as q."
        assert attn_mask.device == q.device, "All tensors must be on the same device."
        assert not is_causal, "Mask should be None for causal attention."
    _tensor_layout = 0 if tensor_layout == "NHD" else 1
    q, k, v, head_dim_og = _pad_head_dim(q, k, v)
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of qkv must be contiguous."
    lse_correction = None
    if smooth_k:
        km = k_mean(k, tensor_layout)
        if return_lse:
            lse_correction = _lse_correction(q, km, tensor_layout)
    else:
        km = None
    if sm_scale is None:
        sm_scale = 1.0 / (head_dim_og ** 0.5)
    if quantization_backend not in ("triton", "cuda"):
        raise ValueError(f"Unsupported quantization backend: {quantization_backend}")
    q_int8, q_scale, k_int8, k_scale = per_block_int8(q, k, km=km, sm_scale=sm_scale, tensor_layout=tensor_layout,
                                                      semantics=quantization_backend)
    v_t = transpose_v_f16(v, tensor_layout=tensor_layout)       # `v.to(torch.float16)`, core.py:297-298
    o = torch.empty(q.size(), dtype=dtype, device=q.device)
    if attn_mask is not None:
        if tensor_layout == "HND":
            target_shape = (q.shape[0], q.shape[1], q.shape[2], k.shape[2])
        else:
            target_shape = (q.shape[0], q.shape[2], q.shape[1], k.shape[1])
        try:
            attn_mask = attn_mask.expand(target_shape)     # core.py:314-323; a view: broadcast dims get stride 0
        except Exception:
            raise AssertionError(f"attn_mask shape {attn_mask.shape} cannot be broadcast to {target_shape}")
        lse = ops.qk_int8_sv_f16_attn_masked(q_int8, k_int8, v_t, o, q_scale, k_scale, attn_mask, _tensor_layout,
                                             SAB_GRAN_PER_BLOCK, SAB_GRAN_PER_BLOCK, sm_scale, 1, 1 if return_lse else 0)
    else:
        lse = ops.qk_int8_sv_f16_attn(q_int8, k_int8, v_t, o, q_scale, k_scale, _tensor_layout, 1 if is_causal else 0,
                                      SAB_GRAN_PER_BLOCK, SAB_GRAN_PER_BLOCK, sm_scale, 1, 1 if return_lse else 0)
    o = o[..., :head_dim_og]
    if return_lse:
        return o, lse / _LOG2E + lse_correction * sm_scale if smooth_k else lse / _LOG2E
    return o


def sageattn_qk_int8_pv_fp16_cuda(q, k, v, tensor_layout: str = "HND", is_causal: bool = False,
                                  qk_quant_gran: str = "per_thread", sm_scale: Optional[float] = None,
                                  pv_accum_dtype: str = "fp32", smooth_k: bool = True, smooth_v: bool = False,
                                  return_lse: bool = False, **kwargs: Any):
    """sageattention/core.py:451-633 (INT8 QK^T + **FP16** PV) on sm_100a: V and P stay in fp16 — callers pick this entry
    over the fp8 ones for accuracy — through the FP16-PV kernel variant (tcgen05 kind::f16, fp32 accumulation in TMEM).
    Q/K quantisation is the reference's per-warp / per-thread INT8 (core.py:590-593; WARPQ is 32 for every mode here: the
    scale packing is internal to this call).  pv_accum_dtype keeps the reference vocabulary ("fp32", "fp16", "fp16+fp32",
    core.py:601-617): the B200 tensor core always accumulates PV in fp32, i.e. every mode gets the accuracy of "fp32".
    smooth_v exists in the reference only to protect the "fp16" accumulator from overflow (core.py:606-610); with fp32
    accumulation it has nothing to do, so it is accepted and ignored with the reference's warning."""
    dtype = q.dtype
    _check_inputs(q, k, v)
    assert qk_quant_gran in ["per_warp", "per_thread"], "qk_quant_gran must be either 'per_warp' or 'per_thread'."
    if pv_accum_dtype not in ("fp32", "fp16", "fp16+fp32"):
        raise ValueError(f"Unsupported pv_accum_dtype: {pv_accum_dtype}")
    if tensor_layout not in ("NHD", "HND"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    _tensor_layout = 0 if tensor_layout == "NHD" else 1
    q, k, v, head_dim_og = _pad_head_dim(q, k, v)
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of qkv must be contiguous."
    if is_causal:
        seq_dim = 1 if _tensor_layout == 0 else 2
        assert q.size(seq_dim) == k.size(seq_dim), "qo_len and kv_len must be equal for causal attention."
    if sm_scale is None:
        sm_scale = head_dim_og ** -0.5
    lse_correction = None
    km = k_mean(k, tensor_layout) if smooth_k else None
    if smooth_k and return_lse:
        lse_correction = _lse_correction(q, km, tensor_layout)
    q_int8, q_scale = quant_q_int8(q, qk_quant_gran, tensor_layout)
    k_int8, k_scale = quant_k_int8(k, km, qk_quant_gran, tensor_layout)
    gran = SAB_GRAN_PER_WARP if qk_quant_gran == "per_warp" else SAB_GRAN_PER_THREAD
    if smooth_v:
        warnings.warn(f"pv_accum_dtype is '{pv_accum_dtype}' (fp32 accumulation on B200), smooth_v will be ignored.")
    v_t = transpose_v_f16(v, tensor_layout=tensor_layout)       # `v.to(torch.float16)`, core.py:603
    o = torch.empty(q.size(), dtype=dtype, device=q.device)
    lse = ops.qk_int8_sv_f16_attn(q_int8, k_int8, v_t, o, q_scale, k_scale, _tensor_layout, 1 if is_causal else 0,
                                  gran, gran, sm_scale, 0, 1 if return_lse else 0)
    o = o[..., :head_dim_og]
    if return_lse:
        return o, lse / _LOG2E + lse_correction * sm_scale if smooth_k else lse / _LOG2E
    return o


def sageattn_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens_q: torch.Tensor,
                    cu_seqlens_k: torch.Tensor, max_seqlen_q: int, max_seqlen_k: int, is_causal: bool = False,
                    sm_scale: Optional[float] = None, smooth_k: bool = True, **kwargs: Any) -> torch.Tensor:
    """sageattention/core.py:334-448: packed [T,H,D] tensors, per-block INT8 per sequence (Triton rounding,
    bit-exact), K mean over all tokens of the batch (core.py:433), FP16 P and V (tcgen05 kind::f16) like the
    reference's Triton varlen kernels."""
    dtype = q.dtype
    _check_inputs(q, k, v)
    q, k, v, head_dim_og = _pad_head_dim(q, k, v)
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Last dim of qkv must be contiguous."
    assert cu_seqlens_q.is_contiguous() and cu_seqlens_k.is_contiguous(), "cu_seqlens_q and cu_seqlens_k must be contiguous."
    km = None
    if smooth_k:
        km = k_mean(k.unsqueeze(0), "NHD").view(1, k.size(1), k.size(2))   # mean over all tokens
    if sm_scale is None:
        sm_scale = 1.0 / (head_dim_og ** 0.5)
    q_int8, q_scale, k_int8, k_scale, cu_q_scale, cu_k_scale = per_block_int8_varlen(
        q, k, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, sm_scale=sm_scale, km=km)
    v_t, cu_pad = transpose_v_f16_varlen(v, cu_seqlens_k, max_seqlen_k)
    o = torch.empty(q.shape, dtype=dtype, device=q.device)
    ops.qk_int8_sv_f16_attn_varlen(q_int8, k_int8, v_t, o, q_scale, k_scale, cu_seqlens_q.to(torch.int32),
                                   cu_seqlens_k.to(torch.int32), cu_pad, cu_q_scale, cu_k_scale, max_seqlen_q,
                                   1 if is_causal else 0, sm_scale, 1)
    return o[..., :head_dim_og]

"""torch.library custom-op shims over the C ABI (mirrors sageattention/sm89_compile.py:5-146 of the
reference: ops mutate pre-allocated outputs and have fake impls so they trace under torch.compile).

PyTorch is plumbing here: it owns device memory and the current stream; every op body is a single
call into libsageattn_b200.so through ctypes with raw pointers, sizes and strides.
"""
from typing import Optional
import torch

from . import _capi
from ._capi import check, lib, SAB_MASK_BOOL, SAB_MASK_BIAS


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _capi.SAB_DTYPE_FP16
    if t.dtype == torch.bfloat16:
        return _capi.SAB_DTYPE_BF16
    raise TypeError("Only half and bfloat16 are supported")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _bhs_strides(t: torch.Tensor, tensor_layout: int):
    """(stride_b, stride_h, stride_s) in elements; tensor_layout 0 = NHD, 1 = HND (reference encoding)."""
    if tensor_layout == 1:
        return t.stride(0), t.stride(1), t.stride(2)
    return t.stride(0), t.stride(2), t.stride(1)


def _bhsd(t: torch.Tensor, tensor_layout: int):
    if tensor_layout == 1:
        return t.size(0), t.size(1), t.size(2), t.size(3)
    return t.size(0), t.size(2), t.size(1), t.size(3)


# ----------------------------------------------------------------------------------------------- K mean
@torch.library.custom_op("sageattention_b200::k_mean", mutates_args=("mean",), device_types="cuda")
def k_mean(k: torch.Tensor, mean: torch.Tensor, tensor_layout: int) -> None:
    B, H, S, D = _bhsd(k, tensor_layout)
    sb, sh, ss = _bhs_strides(k, tensor_layout)
    with torch.cuda.device(k.device):
        ws = torch.empty(lib().sab_k_mean_workspace_bytes(B, H, S, D), dtype=torch.uint8, device=k.device)
        check(lib().sab_k_mean(k.data_ptr(), _dt(k), mean.data_ptr(), B, H, S, D, sb, sh, ss, ws.data_ptr(), _stream(k)))


@k_mean.register_fake
def _(k, mean, tensor_layout):
    return None


# ----------------------------------------------------------------------------------------------- fused single-pass front-end
@torch.library.custom_op("sageattention_b200::k_smooth_quant_int8", mutates_args=("mean", "output", "scale"), device_types="cuda")
def k_smooth_quant_int8(k: torch.Tensor, mean: torch.Tensor, output: torch.Tensor, scale: torch.Tensor, tensor_layout: int,
                        granularity: int) -> None:
    """K smoothing mean + INT8 quantisation of K in one launch (cluster per (b,h)); mean [B,H,D] in k.dtype is written too."""
    B, H, S, D = _bhsd(k, tensor_layout)
    xs, os_ = _bhs_strides(k, tensor_layout), _bhs_strides(output, tensor_layout)
    with torch.cuda.device(k.device):
        check(lib().sab_k_smooth_quant_int8(k.data_ptr(), _dt(k), mean.data_ptr(), output.data_ptr(), scale.data_ptr(), B, H, S, D,
                                            *xs, *os_, scale.size(-1), granularity, _stream(k)))


@k_smooth_quant_int8.register_fake
def _(k, mean, output, scale, tensor_layout, granularity):
    return None


@torch.library.custom_op("sageattention_b200::per_channel_fp8_fused", mutates_args=("v_fp8", "v_scale", "v_mean"), device_types="cuda")
def per_channel_fp8_fused(v: torch.Tensor, v_fp8: torch.Tensor, v_scale: torch.Tensor, v_mean: Optional[torch.Tensor],
                          tensor_layout: int, scale_max: float) -> None:
    """per_channel_fp8 (dense form) in one launch: per-channel statistics through a cluster reduction, then quantise + transpose."""
    B, H, S, D = _bhsd(v, tensor_layout)
    sb, sh, ss = _bhs_strides(v, tensor_layout)
    with torch.cuda.device(v.device):
        check(lib().sab_per_channel_fp8_fused(v.data_ptr(), _dt(v), v_fp8.data_ptr(), v_scale.data_ptr(), _ptr(v_mean), B, H, S, D,
                                              sb, sh, ss, v_fp8.size(-1), float(scale_max), _stream(v)))


@per_channel_fp8_fused.register_fake
def _(v, v_fp8, v_scale, v_mean, tensor_layout, scale_max):
    return None


# ----------------------------------------------------------------------------------------------- INT8 quant
@torch.library.custom_op("sageattention_b200::quant_per_block_int8", mutates_args=("output", "scale"), device_types="cuda")
def quant_per_block_int8(input: torch.Tensor, mean: Optional[torch.Tensor], output: torch.Tensor, scale: torch.Tensor,
                         block_size: int, tensor_layout: int, semantics: int, has_sm_scale: bool, sm_scale: float) -> None:
    B, H, S, D = _bhsd(input, tensor_layout)
    xs, os_ = _bhs_strides(input, tensor_layout), _bhs_strides(output, tensor_layout)
    with torch.cuda.device(input.device):
        check(lib().sab_quant_per_block_int8(input.data_ptr(), _dt(input), _ptr(mean), output.data_ptr(), scale.data_ptr(),
                                             B, H, S, D, *xs, *os_, scale.size(-1), block_size, semantics,
                                             1 if has_sm_scale else 0, float(sm_scale), _stream(input)))


@quant_per_block_int8.register_fake
def _(input, mean, output, scale, block_size, tensor_layout, semantics, has_sm_scale, sm_scale):
    return None


@torch.library.custom_op("sageattention_b200::quant_per_thread_int8", mutates_args=("output", "scale"), device_types="cuda")
def quant_per_thread_int8(input: torch.Tensor, mean: Optional[torch.Tensor], output: torch.Tensor, scale: torch.Tensor,
                          tensor_layout: int, is_key: bool) -> None:
    B, H, S, D = _bhsd(input, tensor_layout)
    xs, os_ = _bhs_strides(input, tensor_layout), _bhs_strides(output, tensor_layout)
    with torch.cuda.device(input.device):
        check(lib().sab_quant_per_thread_int8(input.data_ptr(), _dt(input), _ptr(mean), output.data_ptr(), scale.data_ptr(),
                                              B, H, S, D, *xs, *os_, scale.size(-1), 1 if is_key else 0, _stream(input)))


@quant_per_thread_int8.register_fake
def _(input, mean, output, scale, tensor_layout, is_key):
    return None


@torch.library.custom_op("sageattention_b200::quant_per_block_int8_varlen", mutates_args=("output", "scale"), device_types="cuda")
def quant_per_block_int8_varlen(input: torch.Tensor, mean: Optional[torch.Tensor], output: torch.Tensor, scale: torch.Tensor,
                                cu_seqlens: torch.Tensor, cu_scale: torch.Tensor, max_seqlen: int, block_size: int,
                                has_sm_scale: bool, sm_scale: float) -> None:
    T, H, D = input.shape
    with torch.cuda.device(input.device):
        check(lib().sab_quant_per_block_int8_varlen(input.data_ptr(), _dt(input), _ptr(mean), output.data_ptr(), scale.data_ptr(),
                                                    cu_seqlens.data_ptr(), cu_scale.data_ptr(), cu_seqlens.numel() - 1,
                                                    max_seqlen, H, D, input.stride(0), input.stride(1), output.stride(0),
                                                    output.stride(1), block_size, 1 if has_sm_scale else 0, float(sm_scale),
                                                    _stream(input)))


@quant_per_block_int8_varlen.register_fake
def _(input, mean, output, scale, cu_seqlens, cu_scale, max_seqlen, block_size, has_sm_scale, sm_scale):
    return None


# ----------------------------------------------------------------------------------------------- FP8 V
@torch.library.custom_op("sageattention_b200::per_channel_fp8", mutates_args=("v_fp8", "v_scale", "v_mean"), device_types="cuda")
def per_channel_fp8(v: torch.Tensor, v_fp8: torch.Tensor, v_scale: torch.Tensor, v_mean: Optional[torch.Tensor],
                    tensor_layout: int, scale_max: float) -> None:
    B, H, S, D = _bhsd(v, tensor_layout)
    sb, sh, ss = _bhs_strides(v, tensor_layout)
    with torch.cuda.device(v.device):
        ws = torch.empty(lib().sab_per_channel_fp8_workspace_bytes(B, H, S, D), dtype=torch.uint8, device=v.device)
        check(lib().sab_per_channel_fp8(v.data_ptr(), _dt(v), v_fp8.data_ptr(), v_scale.data_ptr(), _ptr(v_mean), B, H, S, D,
                                        sb, sh, ss, v_fp8.size(-1), float(scale_max), None, None, 0, 0, ws.data_ptr(), _stream(v)))


@per_channel_fp8.register_fake
def _(v, v_fp8, v_scale, v_mean, tensor_layout, scale_max):
    return None


@torch.library.custom_op("sageattention_b200::per_channel_fp8_varlen", mutates_args=("v_fp8", "v_scale"), device_types="cuda")
def per_channel_fp8_varlen(v: torch.Tensor, v_fp8: torch.Tensor, v_scale: torch.Tensor, cu_seqlens: torch.Tensor,
                           cu_pad: torch.Tensor, max_seqlen: int, scale_max: float) -> None:
    T, H, D = v.shape
    with torch.cuda.device(v.device):
        ws = torch.empty(lib().sab_per_channel_fp8_workspace_bytes(1, H, T, D), dtype=torch.uint8, device=v.device)
        check(lib().sab_per_channel_fp8(v.data_ptr(), _dt(v), v_fp8.data_ptr(), v_scale.data_ptr(), None, 1, H, T, D,
                                        0, v.stride(1), v.stride(0), v_fp8.size(-1), float(scale_max), cu_seqlens.data_ptr(),
                                        cu_pad.data_ptr(), cu_seqlens.numel() - 1, max_seqlen, ws.data_ptr(), _stream(v)))


@per_channel_fp8_varlen.register_fake
def _(v, v_fp8, v_scale, cu_seqlens, cu_pad, max_seqlen, scale_max):
    return None


# ----------------------------------------------------------------------------------------------- SP building blocks
@torch.library.custom_op("sageattention_b200::channel_stats", mutates_args=("sum_out", "max_out", "min_out"), device_types="cuda")
def channel_stats(x: torch.Tensor, sum_out: torch.Tensor, max_out: torch.Tensor, min_out: torch.Tensor, tensor_layout: int) -> None:
    B, H, S, D = _bhsd(x, tensor_layout)
    sb, sh, ss = _bhs_strides(x, tensor_layout)
    with torch.cuda.device(x.device):
        ws = torch.empty(lib().sab_k_mean_workspace_bytes(B, H, S, D), dtype=torch.uint8, device=x.device)
        check(lib().sab_channel_stats(x.data_ptr(), _dt(x), sum_out.data_ptr(), max_out.data_ptr(), min_out.data_ptr(),
                                      B, H, S, D, sb, sh, ss, ws.data_ptr(), _stream(x)))


@channel_stats.register_fake
def _(x, sum_out, max_out, min_out, tensor_layout):
    return None


@torch.library.custom_op("sageattention_b200::v_quant_with_amax", mutates_args=("v_fp8", "v_scale"), device_types="cuda")
def v_quant_with_amax(v: torch.Tensor, v_fp8: torch.Tensor, amax: torch.Tensor, v_scale: torch.Tensor, tensor_layout: int,
                      scale_max: float) -> None:
    B, H, S, D = _bhsd(v, tensor_layout)
    sb, sh, ss = _bhs_strides(v, tensor_layout)
    with torch.cuda.device(v.device):
        ws = torch.empty(B * H * D * 4, dtype=torch.uint8, device=v.device)
        check(lib().sab_v_quant_with_amax(v.data_ptr(), _dt(v), v_fp8.data_ptr(), amax.data_ptr(), v_scale.data_ptr(), B, H, S, D,
                                          sb, sh, ss, v_fp8.size(-1), float(scale_max), ws.data_ptr(), _stream(v)))


@v_quant_with_amax.register_fake
def _(v, v_fp8, amax, v_scale, tensor_layout, scale_max):
    return None


# ----------------------------------------------------------------------------------------------- attention
@torch.library.custom_op("sageattention_b200::qk_int8_sv_f8_attn", mutates_args=("output",), device_types="cuda")
def qk_int8_sv_f8_attn(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, output: torch.Tensor,
                       query_scale: torch.Tensor, key_scale: torch.Tensor, value_scale: Optional[torch.Tensor],
                       value_mean: Optional[torch.Tensor], tensor_layout: int, is_causal: int, q_quant_gran: int,
                       k_quant_gran: int, sm_scale: float, fold_sm_scale: int, return_lse: int,
                       causal_q_offset: int = 0, kv_seg_len: int = 0) -> torch.Tensor:
    """Same contract as the reference op (sm89_compile.py:48-66): mutates `output`, returns lse
    ([B,Hq,Sq] fp32 in log2 units, or an empty tensor when return_lse == 0).
    kv_seg_len > 0: key/value are the rank-major all-gather of P shards ([P*B,Hkv,kv_seg_len,D] /
    [P*B,Hkv,D,kv_seg_len]); causal_q_offset is the global index of query row 0."""
    B, Hq, Sq, D = _bhsd(query, tensor_layout)
    _, Hkv, Skv, _ = _bhsd(key, tensor_layout)
    if kv_seg_len > 0:
        Skv = (key.size(0) // B) * kv_seg_len
    qs, ks, os_ = _bhs_strides(query, tensor_layout), _bhs_strides(key, tensor_layout), _bhs_strides(output, tensor_layout)
    lse = torch.empty((B, Hq, Sq), dtype=torch.float32, device=query.device) if return_lse else \
        torch.empty((0,), dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        check(lib().sab_qk_int8_sv_f8_attn(query.data_ptr(), key.data_ptr(), value.data_ptr(), output.data_ptr(),
                                           lse.data_ptr() if return_lse else None, query_scale.data_ptr(), key_scale.data_ptr(),
                                           _ptr(value_scale), _ptr(value_mean), _dt(output), B, Hq, Hkv, Sq, Skv, D,
                                           *qs, *ks, value.size(-1), *os_, is_causal, q_quant_gran, k_quant_gran,
                                           float(sm_scale), fold_sm_scale, None, None, None, None, None, 0,
                                           causal_q_offset, kv_seg_len, None, _stream(query)))
    return lse


@qk_int8_sv_f8_attn.register_fake
def _(query, key, value, output, query_scale, key_scale, value_scale, value_mean, tensor_layout, is_causal,
      q_quant_gran, k_quant_gran, sm_scale, fold_sm_scale, return_lse, causal_q_offset=0, kv_seg_len=0):
    B, Hq, Sq, D = _bhsd(query, tensor_layout)
    if return_lse:
        return torch.empty((B, Hq, Sq), dtype=torch.float32, device=query.device)
    return torch.empty((0,), dtype=torch.float32, device=query.device)


@torch.library.custom_op("sageattention_b200::qk_int8_sv_f8_attn_sp", mutates_args=("output",), device_types="cuda")
def qk_int8_sv_f8_attn_sp(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, output: torch.Tensor,
                          query_scale: torch.Tensor, key_scale: torch.Tensor, value_scale: torch.Tensor, q_quant_gran: int,
                          k_quant_gran: int, sm_scale: float, kv_seg_len: int, seg_flags: torch.Tensor, seg_epoch: int,
                          heads_per_flag: int) -> None:
    """Sequence-parallel attention whose K/V segments may still be arriving (include/sageattn_b200.h, sab_qk_int8_sv_f8_attn_sp):
    HND tensors, key [P*B,Hkv,kv_seg_len,D] / value [P*B,Hkv,D,kv_seg_len] rank-major, seg_flags uint32 [Hkv/heads_per_flag * P];
    the kernel waits for seg_flags[group * P + segment] == seg_epoch before the first tile of a segment.  Non-causal."""
    B, Hq, Sq, D = query.shape
    Hkv = key.size(1)
    P = key.size(0) // B
    assert seg_flags.dtype in (torch.int32, torch.uint32) and seg_flags.is_contiguous() and seg_flags.numel() >= (Hkv // heads_per_flag) * P
    qs, ks, os_ = _bhs_strides(query, 1), _bhs_strides(key, 1), _bhs_strides(output, 1)
    with torch.cuda.device(query.device):
        check(lib().sab_qk_int8_sv_f8_attn_sp(query.data_ptr(), key.data_ptr(), value.data_ptr(), output.data_ptr(), None,
                                              query_scale.data_ptr(), key_scale.data_ptr(), value_scale.data_ptr(), _dt(output),
                                              B, Hq, Hkv, Sq, P * kv_seg_len, D, *qs, *ks, value.size(-1), *os_, q_quant_gran,
                                              k_quant_gran, float(sm_scale), kv_seg_len, seg_flags.data_ptr(),
                                              int(seg_epoch) & 0xFFFFFFFF, heads_per_flag, _stream(query)))


@qk_int8_sv_f8_attn_sp.register_fake
def _(query, key, value, output, query_scale, key_scale, value_scale, q_quant_gran, k_quant_gran, sm_scale, kv_seg_len,
      seg_flags, seg_epoch, heads_per_flag):
    return None


@torch.library.custom_op("sageattention_b200::qk_int8_sv_f8_attn_varlen", mutates_args=("output",), device_types="cuda")
def qk_int8_sv_f8_attn_varlen(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, output: torch.Tensor,
                              query_scale: torch.Tensor, key_scale: torch.Tensor, value_scale: torch.Tensor,
                              cu_seqlens_q: torch.Tensor, cu_seqlens_k: torch.Tensor, cu_pad_v: torch.Tensor,
                              cu_q_scale: torch.Tensor, cu_k_scale: torch.Tensor, max_seqlen_q: int, is_causal: int,
                              sm_scale: float, fold_sm_scale: int) -> None:
    Tq, Hq, D = query.shape
    Tk, Hkv, _ = key.shape
    nseq = cu_seqlens_q.numel() - 1
    with torch.cuda.device(query.device):
        check(lib().sab_qk_int8_sv_f8_attn(query.data_ptr(), key.data_ptr(), value.data_ptr(), output.data_ptr(), None,
                                           query_scale.data_ptr(), key_scale.data_ptr(), value_scale.data_ptr(), None,
                                           _dt(output), nseq, Hq, Hkv, Tq, Tk, D,
                                           0, query.stride(1), query.stride(0), 0, key.stride(1), key.stride(0),
                                           value.size(-1), 0, output.stride(1), output.stride(0), is_causal,
                                           _capi.SAB_GRAN_PER_BLOCK, _capi.SAB_GRAN_PER_BLOCK, float(sm_scale), fold_sm_scale,
                                           cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr(), cu_pad_v.data_ptr(),
                                           cu_q_scale.data_ptr(), cu_k_scale.data_ptr(), max_seqlen_q, 0, 0, None, _stream(query)))


@qk_int8_sv_f8_attn_varlen.register_fake
def _(query, key, value, output, query_scale, key_scale, value_scale, cu_seqlens_q, cu_seqlens_k, cu_pad_v,
      cu_q_scale, cu_k_scale, max_seqlen_q, is_causal, sm_scale, fold_sm_scale):
    return None


# ----------------------------------------------------------------------------------------------- FP16-PV variant
@torch.library.custom_op("sageattention_b200::v_transpose_f16", mutates_args=("v_f16t",), device_types="cuda")
def v_transpose_f16(v: torch.Tensor, v_f16t: torch.Tensor, tensor_layout: int) -> None:
    B, H, S, D = _bhsd(v, tensor_layout)
    sb, sh, ss = _bhs_strides(v, tensor_layout)
    with torch.cuda.device(v.device):
        check(lib().sab_v_transpose_f16(v.data_ptr(), _dt(v), v_f16t.data_ptr(), B, H, S, D, sb, sh, ss, v_f16t.size(-1),
                                        None, None, 0, 0, _stream(v)))


@v_transpose_f16.register_fake
def _(v, v_f16t, tensor_layout):
    return None


@torch.library.custom_op("sageattention_b200::v_transpose_f16_varlen", mutates_args=("v_f16t",), device_types="cuda")
def v_transpose_f16_varlen(v: torch.Tensor, v_f16t: torch.Tensor, cu_seqlens: torch.Tensor, cu_pad: torch.Tensor,
                           max_seqlen: int) -> None:
    T, H, D = v.shape
    with torch.cuda.device(v.device):
        check(lib().sab_v_transpose_f16(v.data_ptr(), _dt(v), v_f16t.data_ptr(), 1, H, T, D, 0, v.stride(1), v.stride(0),
                                        v_f16t.size(-1), cu_seqlens.data_ptr(), cu_pad.data_ptr(), cu_seqlens.numel() - 1,
                                        max_seqlen, _stream(v)))


@v_transpose_f16_varlen.register_fake
def _(v, v_f16t, cu_seqlens, cu_pad, max_seqlen):
    return None


@torch.library.custom_op("sageattention_b200::qk_int8_sv_f16_attn", mutates_args=("output",), device_types="cuda")
def qk_int8_sv_f16_attn(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, output: torch.Tensor,
                        query_scale: torch.Tensor, key_scale: torch.Tensor, tensor_layout: int, is_causal: int,
                        q_quant_gran: int, k_quant_gran: int, sm_scale: float, fold_sm_scale: int,
                        return_lse: int) -> torch.Tensor:
    """INT8 QK^T + FP16 PV (the reference Triton kernels' numerics); value = [B,Hkv,D,S_pad] fp16 from v_transpose_f16."""
    B, Hq, Sq, D = _bhsd(query, tensor_layout)
    _, Hkv, Skv, _ = _bhsd(key, tensor_layout)
    qs, ks, os_ = _bhs_strides(query, tensor_layout), _bhs_strides(key, tensor_layout), _bhs_strides(output, tensor_layout)
    lse = torch.empty((B, Hq, Sq), dtype=torch.float32, device=query.device) if return_lse else \
        torch.empty((0,), dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        check(lib().sab_qk_int8_sv_f16_attn(query.data_ptr(), key.data_ptr(), value.data_ptr(), output.data_ptr(),
                                            lse.data_ptr() if return_lse else None, query_scale.data_ptr(), key_scale.data_ptr(),
                                            _dt(output), B, Hq, Hkv, Sq, Skv, D, *qs, *ks, value.size(-1), *os_, is_causal,
                                            q_quant_gran, k_quant_gran, float(sm_scale), fold_sm_scale, None, None, None, None,
                                            None, 0, _stream(query)))
    return lse


@qk_int8_sv_f16_attn.register_fake
def _(query, key, value, output, query_scale, key_scale, tensor_layout, is_causal, q_quant_gran, k_quant_gran, sm_scale,
      fold_sm_scale, return_lse):
    B, Hq, Sq, D = _bhsd(query, tensor_layout)
    if return_lse:
        return torch.empty((B, Hq, Sq), dtype=torch.float32, device=query.device)
    return torch.empty((0,), dtype=torch.float32, device=query.device)


@torch.library.custom_op("sageattention_b200::qk_int8_sv_f16_attn_masked", mutates_args=("output",), device_types="cuda")
def qk_int8_sv_f16_attn_masked(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, output: torch.Tensor,
                               query_scale: torch.Tensor, key_scale: torch.Tensor, attn_mask: torch.Tensor, tensor_layout: int,
                               q_quant_gran: int, k_quant_gran: int, sm_scale: float, fold_sm_scale: int,
                               return_lse: int) -> torch.Tensor:
    """qk_int8_sv_f16_attn with the Triton path's attn_mask: a 4-D view broadcast to [B,Hq,Sq,Skv] (strides may be 0),
    dtype bool (False = masked out) or the q dtype (additive bias).  Non-causal (sageattention/core.py:310)."""
    B, Hq, Sq, D = _bhsd(query, tensor_layout)
    _, Hkv, Skv, _ = _bhsd(key, tensor_layout)
    assert attn_mask.dim() == 4 and tuple(attn_mask.shape) == (B, Hq, Sq, Skv), "attn_mask must be expanded to [B,Hq,Sq,Skv]"
    kind = SAB_MASK_BOOL if attn_mask.dtype == torch.bool else SAB_MASK_BIAS
    assert kind == SAB_MASK_BOOL or attn_mask.dtype == output.dtype, "attn_mask must be bool or the q / output dtype"
    qs, ks, os_ = _bhs_strides(query, tensor_layout), _bhs_strides(key, tensor_layout), _bhs_strides(output, tensor_layout)
    lse = torch.empty((B, Hq, Sq), dtype=torch.float32, device=query.device) if return_lse else \
        torch.empty((0,), dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        check(lib().sab_qk_int8_sv_f16_attn_masked(query.data_ptr(), key.data_ptr(), value.data_ptr(), output.data_ptr(),
                                                   lse.data_ptr() if return_lse else None, query_scale.data_ptr(),
                                                   key_scale.data_ptr(), _dt(output), B, Hq, Hkv, Sq, Skv, D, *qs, *ks,
                                                   value.size(-1), *os_, q_quant_gran, k_quant_gran, float(sm_scale),
                                                   fold_sm_scale, attn_mask.data_ptr(), kind, *attn_mask.stride(),
                                                   _stream(query)))
    return lse


@qk_int8_sv_f16_attn_masked.register_fake
def _(query, key, value, output, query_scale, key_scale, attn_mask, tensor_layout, q_quant_gran, k_quant_gran, sm_scale,
      fold_sm_scale, return_lse):
    B, Hq, Sq, D = _bhsd(query, tensor_layout)
    if return_lse:
        return torch.empty((B, Hq, Sq), dtype=torch.float32, device=query.device)
    return torch.empty((0,), dtype=torch.float32, device=query.device)


@torch.library.custom_op("sageattention_b200::qk_int8_sv_f16_attn_varlen", mutates_args=("output",), device_types="cuda")
def qk_int8_sv_f16_attn_varlen(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, output: torch.Tensor,
                               query_scale: torch.Tensor, key_scale: torch.Tensor, cu_seqlens_q: torch.Tensor,
                               cu_seqlens_k: torch.Tensor, cu_pad_v: torch.Tensor, cu_q_scale: torch.Tensor,
                               cu_k_scale: torch.Tensor, max_seqlen_q: int, is_causal: int, sm_scale: float,
                               fold_sm_scale: int) -> None:
    Tq, Hq, D = query.shape
    Tk, Hkv, _ = key.shape
    nseq = cu_seqlens_q.numel() - 1
    with torch.cuda.device(query.device):
        check(lib().sab_qk_int8_sv_f16_attn(query.data_ptr(), key.data_ptr(), value.data_ptr(), output.data_ptr(), None,
                                            query_scale.data_ptr(), key_scale.data_ptr(), _dt(output), nseq, Hq, Hkv, Tq, Tk, D,
                                            0, query.stride(1), query.stride(0), 0, key.stride(1), key.stride(0), value.size(-1),
                                            0, output.stride(1), output.stride(0), is_causal, _capi.SAB_GRAN_PER_BLOCK,
                                            _capi.SAB_GRAN_PER_BLOCK, float(sm_scale), fold_sm_scale, cu_seqlens_q.data_ptr(),
                                            cu_seqlens_k.data_ptr(), cu_pad_v.data_ptr(), cu_q_scale.data_ptr(),
                                            cu_k_scale.data_ptr(), max_seqlen_q, _stream(query)))


@qk_int8_sv_f16_attn_varlen.register_fake
def _(query, key, value, output, query_scale, key_scale, cu_seqlens_q, cu_seqlens_k, cu_pad_v, cu_q_scale, cu_k_scale,
      max_seqlen_q, is_causal, sm_scale, fold_sm_scale):
    return None

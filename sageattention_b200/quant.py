"""Quantisation front-end with the reference's Python signatures (sageattention/quant.py:22-293,
sageattention/triton/quant_per_thread.py:154-203, quant_per_block.py:49-101,
quant_per_block_varlen.py:60-104) on top of the sm_100a kernels.  Outputs are allocated here and
passed to the C ABI (the reference's ownership convention)."""
from typing import Optional, Tuple
import torch

from . import ops
from ._capi import SAB_SEM_CUDA, SAB_SEM_TRITON

_LOG2E = 1.44269504


def _layout(tensor_layout: str) -> int:
    if tensor_layout == "HND":
        return 1
    if tensor_layout == "NHD":
        return 0
    raise ValueError(f"Unknown tensor layout: {tensor_layout}")


def _dims(t: torch.Tensor, tensor_layout: str):
    if tensor_layout == "HND":
        b, h, s, d = t.shape
    elif tensor_layout == "NHD":
        b, s, h, d = t.shape
    else:
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    return b, h, s, d


def _squeeze_mean(km: Optional[torch.Tensor], b: int, h: int, d: int) -> Optional[torch.Tensor]:
    """Accept km as [B,H,D] or with the reference's keepdim seq axis ([B,H,1,D] / [B,1,H,D])."""
    if km is None:
        return None
    km = km.reshape(b, h, d) if km.numel() == b * h * d else km
    assert km.shape == (b, h, d), f"km must have {b * h * d} elements"
    return km.contiguous()


def k_mean(k: torch.Tensor, tensor_layout: str = "HND") -> torch.Tensor:
    """`k.mean(dim=seq, keepdim=True)` (sageattention/core.py:773): fp32 accumulation, result in k.dtype,
    returned with the reference's keepdim shape."""
    b, h, s, d = _dims(k, tensor_layout)
    mean = torch.empty((b, h, d), dtype=k.dtype, device=k.device)
    ops.k_mean(k, mean, _layout(tensor_layout))
    return mean.view(b, h, 1, d) if tensor_layout == "HND" else mean.view(b, 1, h, d)


def per_block_int8(q: torch.Tensor, k: torch.Tensor, km: Optional[torch.Tensor] = None, BLKQ: int = 128,
                   BLKK: int = 64, sm_scale: Optional[float] = None, tensor_layout: str = "HND",
                   semantics: str = "triton") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-block INT8 of q (pre-scaled by sm_scale*log2e) and k (optionally smoothed by km).
    semantics="triton" reproduces sageattention/triton/quant_per_block.py:49-101 bit-exactly;
    semantics="cuda" reproduces sageattention/quant.py:22-103 (csrc/fused/fused.cu)."""
    b, h_qo, qo_len, head_dim = _dims(q, tensor_layout)
    _, h_kv, kv_len, _ = _dims(k, tensor_layout)
    lay = _layout(tensor_layout)
    sem = SAB_SEM_TRITON if semantics == "triton" else SAB_SEM_CUDA
    q_int8 = torch.empty(q.shape, dtype=torch.int8, device=q.device)
    k_int8 = torch.empty(k.shape, dtype=torch.int8, device=k.device)
    q_scale = torch.empty((b, h_qo, (qo_len + BLKQ - 1) // BLKQ), device=q.device, dtype=torch.float32)
    k_scale = torch.empty((b, h_kv, (kv_len + BLKK - 1) // BLKK), device=q.device, dtype=torch.float32)
    if sm_scale is None:
        sm_scale = head_dim ** -0.5
    ops.quant_per_block_int8(q, None, q_int8, q_scale, BLKQ, lay, sem, True, sm_scale * _LOG2E)
    ops.quant_per_block_int8(k, _squeeze_mean(km, b, h_kv, head_dim), k_int8, k_scale, BLKK, lay, sem, False, 1.0)
    return q_int8, q_scale, k_int8, k_scale


def quant_q_int8(q: torch.Tensor, qk_quant_gran: str = "per_thread", tensor_layout: str = "HND"):
    """Q side of per_warp_int8 / per_thread_int8 (BLKQ=128, WARPQ=32): returns (q_int8, q_scale)."""
    b, h_qo, qo_len, _ = _dims(q, tensor_layout)
    lay = _layout(tensor_layout)
    q_int8 = torch.empty(q.shape, dtype=torch.int8, device=q.device)
    nblk = (qo_len + 127) // 128 * 4
    if qk_quant_gran == "per_warp":
        q_scale = torch.empty((b, h_qo, nblk), device=q.device, dtype=torch.float32)
        ops.quant_per_block_int8(q, None, q_int8, q_scale, 32, lay, SAB_SEM_CUDA, False, 1.0)
    else:
        q_scale = torch.empty((b, h_qo, nblk * 8), device=q.device, dtype=torch.float32)
        ops.quant_per_thread_int8(q, None, q_int8, q_scale, lay, False)
    return q_int8, q_scale


def quant_k_int8(k: torch.Tensor, km: Optional[torch.Tensor] = None, qk_quant_gran: str = "per_thread",
                 tensor_layout: str = "HND"):
    """K side of per_warp_int8 / per_thread_int8 (BLKK=64, WARPK=64) with the fused `k - km`: returns (k_int8, k_scale)."""
    b, h_kv, kv_len, head_dim = _dims(k, tensor_layout)
    lay = _layout(tensor_layout)
    k_int8 = torch.empty(k.shape, dtype=torch.int8, device=k.device)
    nblk = (kv_len + 63) // 64
    mean = _squeeze_mean(km, b, h_kv, head_dim)
    if qk_quant_gran == "per_warp":
        k_scale = torch.empty((b, h_kv, nblk), device=k.device, dtype=torch.float32)
        ops.quant_per_block_int8(k, mean, k_int8, k_scale, 64, lay, SAB_SEM_CUDA, False, 1.0)
    else:
        k_scale = torch.empty((b, h_kv, nblk * 4), device=k.device, dtype=torch.float32)
        ops.quant_per_thread_int8(k, mean, k_int8, k_scale, lay, True)
    return k_int8, k_scale


def smooth_quant_k(k: torch.Tensor, qk_quant_gran: str = "per_thread", tensor_layout: str = "HND"):
    """`km = k.mean(dim=seq, keepdim=True)` (sageattention/core.py:773) + the K half of per_thread_int8 / per_warp_int8 in ONE launch
    (csrc/quant.cu k_smooth_quant_kernel: a thread-block cluster per (b,h) reduces the mean through distributed shared memory and
    quantises right away, so K crosses HBM once).  Returns (km, k_int8, k_scale); k_int8 / k_scale are bit-identical to
    quant_k_int8(k, km) with the returned km (km itself is a different fp32 summation order than k_mean(): same value up to one
    ulp on rare channels).  OPT-IN: measured on B200 the one launch is slower than k_mean + quant_k_int8 (177 vs 153 us at
    4x32x8192x128) — the quantisers are latency- not bandwidth-bound, so saving the second HBM read does not pay."""
    from ._capi import SAB_GRAN_PER_WARP, SAB_GRAN_PER_THREAD
    b, h_kv, kv_len, head_dim = _dims(k, tensor_layout)
    lay = _layout(tensor_layout)
    k_int8 = torch.empty(k.shape, dtype=torch.int8, device=k.device)
    nblk = (kv_len + 63) // 64
    mean = torch.empty((b, h_kv, head_dim), dtype=k.dtype, device=k.device)
    if qk_quant_gran == "per_warp":
        k_scale = torch.empty((b, h_kv, nblk), device=k.device, dtype=torch.float32)
        gran = SAB_GRAN_PER_WARP
    else:
        k_scale = torch.empty((b, h_kv, nblk * 4), device=k.device, dtype=torch.float32)
        gran = SAB_GRAN_PER_THREAD
    ops.k_smooth_quant_int8(k, mean, k_int8, k_scale, lay, gran)
    km = mean.view(b, h_kv, 1, head_dim) if tensor_layout == "HND" else mean.view(b, 1, h_kv, head_dim)
    return km, k_int8, k_scale


def per_warp_int8(q: torch.Tensor, k: torch.Tensor, km: Optional[torch.Tensor] = None, BLKQ: int = 128,
                  WARPQ: int = 32, BLKK: int = 64, tensor_layout: str = "HND"):
    """sageattention/quant.py:105-180: q per WARPQ-row block, k per BLKK block with fused (k - km) in fp32;
    CUDA rounding semantics (cvt.rni, amax floored at 1e-7)."""
    assert BLKQ == 128 and WARPQ == 32 and BLKK == 64, "sm_100a kernel supports the reference defaults"
    q_int8, q_scale = quant_q_int8(q, "per_warp", tensor_layout)
    k_int8, k_scale = quant_k_int8(k, km, "per_warp", tensor_layout)
    return q_int8, q_scale, k_int8, k_scale


def per_thread_int8(q: torch.Tensor, k: torch.Tensor, km: Optional[torch.Tensor] = None, BLKQ: int = 128,
                    WARPQ: int = 32, BLKK: int = 64, WARPK: int = 64, sm_scale: Optional[float] = None,
                    tensor_layout: str = "HND"):
    """sageattention/triton/quant_per_thread.py:154-203 (8 q scales per 32-row block, 4 k scales per
    64-key block; `k - km` rounded to the input dtype first; scale = amax/127 + 1e-7)."""
    assert BLKQ == 128 and WARPQ == 32 and BLKK == 64 and WARPK == 64, "sm_100a kernel supports the reference defaults"
    q_int8, q_scale = quant_q_int8(q, "per_thread", tensor_layout)
    k_int8, k_scale = quant_k_int8(k, km, "per_thread", tensor_layout)
    return q_int8, q_scale, k_int8, k_scale


def per_block_int8_varlen(q, k, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, BLKQ=128, BLKK=64,
                          sm_scale=None, km: Optional[torch.Tensor] = None):
    """sageattention/triton/quant_per_block_varlen.py:60-104 on packed [T,H,D] tensors.  `km`
    ([1,H,D] or [H,D]) is subtracted inside the kernel (the reference does `k = k - km` in torch first)."""
    h_qo, h_kv, head_dim = q.shape[1], k.shape[1], q.shape[-1]
    q_int8 = torch.empty(q.shape, dtype=torch.int8, device=q.device)
    k_int8 = torch.empty(k.shape, dtype=torch.int8, device=k.device)
    cu_q32, cu_k32 = cu_seqlens_q.to(torch.int32), cu_seqlens_k.to(torch.int32)
    q_batch_len = cu_q32[1:] - cu_q32[:-1]
    k_batch_len = cu_k32[1:] - cu_k32[:-1]
    cu_seqlens_q_scale = torch.nn.functional.pad(torch.cumsum((q_batch_len + BLKQ - 1) // BLKQ, dim=0), (1, 0), value=0).to(torch.int32)
    cu_seqlens_k_scale = torch.nn.functional.pad(torch.cumsum((k_batch_len + BLKK - 1) // BLKK, dim=0), (1, 0), value=0).to(torch.int32)
    nseq = cu_q32.numel() - 1
    # upper bounds avoid a device->host sync (reference: q_scale rows = cu_seqlens_q_scale[-1])
    q_scale = torch.empty((q.shape[0] // BLKQ + nseq, h_qo), device=q.device, dtype=torch.float32)
    k_scale = torch.empty((k.shape[0] // BLKK + nseq, h_kv), device=k.device, dtype=torch.float32)
    if sm_scale is None:
        sm_scale = head_dim ** -0.5
    kmean = None if km is None else km.reshape(h_kv, head_dim).contiguous()
    ops.quant_per_block_int8_varlen(q, None, q_int8, q_scale, cu_q32, cu_seqlens_q_scale, max_seqlen_q, BLKQ, True, sm_scale * _LOG2E)
    ops.quant_per_block_int8_varlen(k, kmean, k_int8, k_scale, cu_k32, cu_seqlens_k_scale, max_seqlen_k, BLKK, False, 1.0)
    return q_int8, q_scale, k_int8, k_scale, cu_seqlens_q_scale, cu_seqlens_k_scale


def per_channel_fp8(v: torch.Tensor, tensor_layout: str = "HND", scale_max: float = 448.0, smooth_v: bool = True, fused: bool = False):
    """Per-channel e4m3 quantisation of V (sageattention/quant.py:224-293).
    Returns (v_fp8, v_scale, vm).  v_fp8 is ``[B, H_kv, D, ceil(kv_len/128)*128]`` float8_e4m3fn
    (token-contiguous "V transposed", zero padded) for BOTH layouts; unlike the reference it is not
    16-token permuted (an mma.sync artefact the tcgen05 kernel does not need)."""
    b, h_kv, kv_len, head_dim = _dims(v, tensor_layout)
    padded_len = (kv_len + 127) // 128 * 128
    v_fp8 = torch.empty((b, h_kv, head_dim, padded_len), dtype=torch.float8_e4m3fn, device=v.device)
    v_scale = torch.empty((b, h_kv, head_dim), dtype=torch.float32, device=v.device)
    vm = torch.empty((b, h_kv, head_dim), dtype=torch.float32, device=v.device) if smooth_v else None
    if fused:   # statistics + quantisation in one cluster launch (measured slower on B200, see DESIGN.md 4.2: opt-in)
        ops.per_channel_fp8_fused(v, v_fp8, v_scale, vm, _layout(tensor_layout), scale_max)
    else:
        ops.per_channel_fp8(v, v_fp8, v_scale, vm, _layout(tensor_layout), scale_max)
    return v_fp8, v_scale, vm


def per_channel_fp8_varlen(v: torch.Tensor, cu_seqlens_k: torch.Tensor, max_seqlen_k: int, scale_max: float = 448.0):
    """Packed-V form: v [T,H,D] -> v_fp8 [H, D, T_pad] where sequence i occupies token columns
    [cu_pad[i], cu_pad[i] + ceil(L_i/128)*128).  Returns (v_fp8, v_scale [H,D], cu_pad)."""
    T, h_kv, head_dim = v.shape
    cu32 = cu_seqlens_k.to(torch.int32)
    lens = cu32[1:] - cu32[:-1]
    cu_pad = torch.nn.functional.pad(torch.cumsum((lens + 127) // 128 * 128, dim=0), (1, 0), value=0).to(torch.int32)
    nseq = cu32.numel() - 1
    t_pad = (T + 127 * nseq + 127) // 128 * 128          # upper bound, no host sync
    v_fp8 = torch.empty((h_kv, head_dim, t_pad), dtype=torch.float8_e4m3fn, device=v.device)
    v_scale = torch.empty((h_kv, head_dim), dtype=torch.float32, device=v.device)
    ops.per_channel_fp8_varlen(v, v_fp8, v_scale, cu32, cu_pad, max_seqlen_k, scale_max)
    return v_fp8, v_scale, cu_pad


def transpose_v_f16(v: torch.Tensor, tensor_layout: str = "HND") -> torch.Tensor:
    """`v.to(torch.float16)` (sageattention/core.py:297-298) in the layout the kind::f16 PV MMA consumes:
    ``[B, H_kv, D, ceil(kv_len/128)*128]`` fp16, token-contiguous, zero padded."""
    b, h_kv, kv_len, head_dim = _dims(v, tensor_layout)
    padded_len = (kv_len + 127) // 128 * 128
    v_t = torch.empty((b, h_kv, head_dim, padded_len), dtype=torch.float16, device=v.device)
    ops.v_transpose_f16(v, v_t, _layout(tensor_layout))
    return v_t


def transpose_v_f16_varlen(v: torch.Tensor, cu_seqlens_k: torch.Tensor, max_seqlen_k: int):
    """Packed form: v [T,H,D] -> [H, D, T_pad] fp16 with every sequence padded to 128 tokens; returns (v_t, cu_pad)."""
    T, h_kv, head_dim = v.shape
    cu32 = cu_seqlens_k.to(torch.int32)
    lens = cu32[1:] - cu32[:-1]
    cu_pad = torch.nn.functional.pad(torch.cumsum((lens + 127) // 128 * 128, dim=0), (1, 0), value=0).to(torch.int32)
    nseq = cu32.numel() - 1
    t_pad = (T + 127 * nseq + 127) // 128 * 128
    v_t = torch.empty((h_kv, head_dim, t_pad), dtype=torch.float16, device=v.device)
    ops.v_transpose_f16_varlen(v, v_t, cu32, cu_pad, max_seqlen_k)
    return v_t, cu_pad

"""Size-independent properties of the CPU oracle (oracle/sage_oracle.py) — the same properties the GPU parity tests use at full
size where the oracle itself is too slow: layout equivalence, GQA = repeated KV heads, causal rows do not see later keys, the K
smoothing makes the result (nearly) invariant to a per-channel key offset, varlen = per-sequence dense calls."""
import numpy as np
import pytest
import torch

from oracle import sage_oracle as O


def _mk(B, H, Hk, Sq, Sk, D, dt=torch.float16, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, Sq, D, generator=g).to(dt)
    k = torch.randn(B, Hk, Sk, D, generator=g).to(dt)
    v = torch.randn(B, Hk, Sk, D, generator=g).to(dt)
    return q, k, v


@pytest.mark.parametrize("gran", ["per_warp", "per_thread"])
def test_nhd_equals_hnd_bit_for_bit(gran):
    q, k, v = _mk(2, 4, 2, 192, 192, 64)
    o1, l1 = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v, qk_quant_gran=gran, return_lse=True, is_causal=True)
    o2, l2 = O.sageattn_qk_int8_pv_fp8_cuda(*(t.transpose(1, 2).contiguous() for t in (q, k, v)), tensor_layout="NHD",
                                            qk_quant_gran=gran, return_lse=True, is_causal=True)
    assert torch.equal(o1, o2.transpose(1, 2)) and torch.equal(l1, l2)


def test_gqa_equals_repeated_kv_heads():
    q, k, v = _mk(1, 6, 2, 256, 256, 128, torch.bfloat16, seed=1)
    o1 = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v)
    o2 = O.sageattn_qk_int8_pv_fp8_cuda(q, k.repeat_interleave(3, 1), v.repeat_interleave(3, 1))
    assert torch.equal(o1, o2)


def test_causal_rows_do_not_depend_on_later_keys():
    """Top-left alignment (attn_utils.cuh:310): the first 128 query rows of a causal call over 384 keys equal the causal call over
    the first 128 keys alone — bit for bit once the per-sequence statistics agree: no smoothing mean, per-warp scales (blocks of
    32 / 64 rows are independent), and a first value row that carries every channel's |max| so that both calls get the same
    per-channel V scale.  Fully masked tiles must contribute exactly nothing (P = 0, alpha = 1)."""
    q, k, v = _mk(1, 2, 2, 384, 384, 64, seed=2)
    v[:, :, 0, :] = 8.0
    kw = dict(is_causal=True, smooth_k=False, qk_quant_gran="per_warp", pv_accum_dtype="fp32+fp32", return_lse=True)
    full, lse_full = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v, **kw)
    head, lse_head = O.sageattn_qk_int8_pv_fp8_cuda(q[:, :, :128], k[:, :, :128], v[:, :, :128], **kw)
    assert torch.equal(full[:, :, :128], head) and torch.equal(lse_full[:, :, :128], lse_head)


def test_smoothing_absorbs_a_channel_offset_of_k():
    """softmax(q (k + c)^T) = softmax(q k^T) for a per-channel offset c; without smoothing the offset wrecks the INT8 range of K,
    with it (core.py:773) the result stays at the un-shifted accuracy."""
    q, k, v = _mk(1, 2, 2, 256, 256, 64, seed=3)
    c = 6.0 * torch.randn(1, 2, 1, 64, generator=torch.Generator().manual_seed(4)).to(k.dtype)
    exact = O.sdpa_fp32(q, k, v)
    base = (O.sageattn_qk_int8_pv_fp8_cuda(q, k, v).float() - exact).abs().mean().item()
    smooth = (O.sageattn_qk_int8_pv_fp8_cuda(q, k + c, v).float() - exact).abs().mean().item()
    rough = (O.sageattn_qk_int8_pv_fp8_cuda(q, k + c, v, smooth_k=False).float() - exact).abs().mean().item()
    assert smooth < 1.1 * base + 1e-4 and rough > 1.3 * smooth


def test_varlen_equals_per_sequence_dense_calls_given_the_batch_mean():
    """sageattn_varlen (core.py:334-448) smooths K with the mean over ALL packed tokens (core.py:433); given that mean, each sequence
    is an independent dense call of the Triton path."""
    lens = [100, 37, 256]
    g = torch.Generator().manual_seed(5)
    T, H, Hk, D = sum(lens), 4, 2, 64
    q = torch.randn(T, H, D, generator=g).half()
    k = (torch.randn(T, Hk, D, generator=g) + 2.0 * torch.randn(1, Hk, D, generator=g)).half()
    v = torch.randn(T, Hk, D, generator=g).half()
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32)
    o = O.sageattn_varlen(q, k, v, cu, cu, max(lens), max(lens))
    km = k.mean(dim=0, keepdim=True)
    for i, L in enumerate(lens):
        s = slice(int(cu[i]), int(cu[i + 1]))
        qi, ki, vi = (t[s].transpose(0, 1).unsqueeze(0) for t in (q, k - km, v))          # [1,H,L,D], K already smoothed
        oi = O.sageattn_qk_int8_pv_fp16_triton(qi, ki, vi, smooth_k=False)
        assert (o[s].transpose(0, 1).unsqueeze(0).float() - oi.float()).abs().max().item() <= 2e-3, i


def test_fp16_pv_cuda_entry_is_more_accurate_than_the_fp8_one_and_respects_causality():
    """sageattn_qk_int8_pv_fp16_cuda (core.py:451-633): same INT8 QK^T as the fp8 entry but fp16 P and V, so its error against
    exact attention is the QK quantisation error alone — smaller than the fp8 entry's; causal prefix property; NHD == HND."""
    q, k, v = _mk(1, 4, 2, 320, 320, 64, seed=7)
    exact = O.sdpa_fp32(q, k, v)
    e16 = (O.sageattn_qk_int8_pv_fp16_cuda(q, k, v).float() - exact).abs().mean().item()
    e8 = (O.sageattn_qk_int8_pv_fp8_cuda(q, k, v).float() - exact).abs().mean().item()
    assert e16 < e8 and e16 < 5e-3
    oc, lse = O.sageattn_qk_int8_pv_fp16_cuda(q, k, v, is_causal=True, return_lse=True)
    assert (oc.float() - O.sdpa_fp32(q, k, v, is_causal=True)).abs().max().item() < 3e-2
    assert oc.shape == q.shape and lse.shape == q.shape[:3]
    on = O.sageattn_qk_int8_pv_fp16_cuda(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), tensor_layout="NHD", is_causal=True)
    assert torch.equal(on.transpose(1, 2), oc)
    ow = O.sageattn_qk_int8_pv_fp16_cuda(q, k, v, qk_quant_gran="per_warp")
    assert (ow.float() - exact).abs().mean().item() < 8e-3

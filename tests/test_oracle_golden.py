"""Pin the CPU oracle (oracle/sage_oracle.py) to fixtures produced by the real reference Triton
kernels (tests/golden/make_golden.py).  CPU only."""
import os
import numpy as np
import pytest
import torch

from oracle import sage_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _t(a, dtype):
    return torch.from_numpy(a.copy()).view(dtype)


def _dtype(s):
    return torch.bfloat16 if "bfloat16" in str(s) else torch.float16


@pytest.mark.parametrize("name", ["quant_d64_fp16", "quant_d128_bf16"])
def test_quant_bit_exact(name):
    z = np.load(f"{G}/{name}.npz")
    dt = _dtype(z["dtype"])
    q, k, km = _t(z["q"], dt), _t(z["k"], dt), _t(z["km"], dt)
    D = q.shape[-1]
    q8, qs, k8, ks = O.per_block_int8_triton(q, k, km=km, sm_scale=D ** -0.5)
    assert np.array_equal(q8.numpy(), z["pb_q8"]) and np.array_equal(k8.numpy(), z["pb_k8"])
    assert np.array_equal(qs.numpy(), z["pb_qs"]) and np.array_equal(ks.numpy(), z["pb_ks"])
    q8, qs, k8, ks = O.quant_per_thread_int8_triton(q, k, km)
    assert np.array_equal(q8.numpy(), z["pt_q8"]) and np.array_equal(k8.numpy(), z["pt_k8"])
    assert np.array_equal(qs.numpy(), z["pt_qs"]) and np.array_equal(ks.numpy(), z["pt_ks"])
    qn, kn = q.transpose(1, 2).contiguous(), k.transpose(1, 2).contiguous()
    q8, qs, k8, ks = O.quant_per_thread_int8_triton(qn, kn, km.transpose(1, 2), tensor_layout="NHD")
    assert np.array_equal(q8.numpy(), z["ptn_q8"]) and np.array_equal(k8.numpy(), z["ptn_k8"])
    assert np.array_equal(qs.numpy(), z["ptn_qs"]) and np.array_equal(ks.numpy(), z["ptn_ks"])


@pytest.mark.parametrize("name", ["attn_d64_fp16_nc", "attn_d64_fp16_c", "attn_d128_fp16_nc_ragged", "attn_xattn_d96_bf16"])
def test_triton_attention_path(name):
    z = np.load(f"{G}/{name}.npz")
    dt = _dtype(z["dtype"])
    q, k, v, o_ref = (_t(z[n], dt) for n in ("q", "k", "v", "o"))
    o, lse = O.sageattn_qk_int8_pv_fp16_triton(q, k, v, is_causal=bool(z["causal"]), return_lse=True)
    # tl.dot(out_dtype=fp16) accumulation order inside the interpreter is not specified: 2e-3 slack; a bf16 output (ulp 2^-7
    # for 1 <= |o| < 2) turns such a difference into one output ulp
    tol = 2e-3 if dt == torch.float16 else 2.0 ** -7 + 1e-6
    assert (o.float() - o_ref.float()).abs().max().item() <= tol
    assert np.allclose(lse.numpy(), z["lse"], atol=2e-4, rtol=1e-5)
    # yard-stick: both are within the reference's own error of exact attention
    exact = O.sdpa_fp32(q, k, v, is_causal=bool(z["causal"]))
    assert (o.float() - exact).abs().max().item() < 5e-2


def _load_mask(z, dt):
    shape = tuple(int(x) for x in z["mask_shape"])
    if str(z["kind"]) == "bool":
        return torch.from_numpy(z["mask"].copy()).view(shape)
    return _t(z["mask"], dt).view(shape)


@pytest.mark.parametrize("name", ["attn_mask_bool_d64", "attn_mask_bias_d128"])
def test_triton_attention_path_with_attn_mask(name):
    """attn_mask of sageattn_qk_int8_pv_fp16_triton: bool (with one all-false block the reference skips, broadcast over
    heads) and additive fp16 bias, against the reference Triton kernel's output."""
    z = np.load(f"{G}/{name}.npz")
    dt = _dtype(z["dtype"])
    q, k, v, o_ref = (_t(z[n], dt) for n in ("q", "k", "v", "o"))
    mask = _load_mask(z, dt)
    o, lse = O.sageattn_qk_int8_pv_fp16_triton(q, k, v, return_lse=True, attn_mask=mask)
    assert (o.float() - o_ref.float()).abs().max().item() < 2e-3
    assert np.allclose(lse.numpy(), z["lse"], atol=2e-4, rtol=1e-5)
    # the mask matters: the unmasked result is far away
    o_nomask = O.sageattn_qk_int8_pv_fp16_triton(q, k, v)
    assert (o_nomask.float() - o_ref.float()).abs().max().item() > 5e-2


@pytest.mark.parametrize("name", ["varlen_gqa_d128_nc", "varlen_gqa_d128_c"])
def test_varlen_path(name):
    z = np.load(f"{G}/{name}.npz")
    q, k, v, o_ref = (_t(z[n], torch.float16) for n in ("q", "k", "v", "o"))
    cu = torch.from_numpy(z["cu"])
    lens = (cu[1:] - cu[:-1]).tolist()
    # quantised tensors + packed scales bit-exact
    km = k.mean(dim=0, keepdim=True)
    q8, qs, cuqs = O.quant_per_block_int8_varlen_triton(q, cu, 128, (1.0 / 128 ** 0.5) * O.LOG2E_PY)
    k8, ks, cuks = O.quant_per_block_int8_varlen_triton(k - km, cu, 64, 1.0)
    assert np.array_equal(q8.numpy(), z["q8"]) and np.array_equal(k8.numpy(), z["k8"])
    assert np.array_equal(qs.numpy(), z["qs"]) and np.array_equal(ks.numpy(), z["ks"])
    assert np.array_equal(cuqs.numpy(), z["cuqs"]) and np.array_equal(cuks.numpy(), z["cuks"])
    o = O.sageattn_varlen(q, k, v, cu, cu, max(lens), max(lens), is_causal=bool(z["causal"]))
    assert (o.float() - o_ref.float()).abs().max().item() < 2e-3


def test_fp8_path_close_to_exact_and_to_triton_path():
    """The CUDA-semantics fp8 restatement has no CPU-runnable reference; sanity-pin it against exact
    attention and against the (golden-pinned) Triton-semantics path on the same inputs."""
    z = np.load(f"{G}/attn_d64_fp16_nc.npz")
    q, k, v = (_t(z[n], torch.float16) for n in ("q", "k", "v"))
    exact = O.sdpa_fp32(q, k, v)
    for gran in ("per_warp", "per_thread"):
        for acc in ("fp32+fp32", "fp32+fp16"):
            o, lse = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v, qk_quant_gran=gran, pv_accum_dtype=acc,
                                                    return_lse=True)
            assert (o.float() - exact).abs().max().item() < 8e-2   # v has a +2 offset: fp8 V rel. error ~2%
            assert np.allclose(lse.numpy(), z["lse"], atol=5e-2)
    oc = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v, is_causal=True)
    assert (oc.float() - O.sdpa_fp32(q, k, v, is_causal=True)).abs().max().item() < 3e-1   # early causal rows copy single V rows: e4m3 V error ~6%

"""`sageattn` for HOST-resident q, k, v: the end-to-end call of a caller whose tensors live in (pinned) host memory.

A user of the reference with host data writes `sageattn(q.cuda(), k.cuda(), v.cuda()).cpu()`: three H2D copies, the
attention, one D2H copy, strictly one after the other — at configs[1] that is 1.07 GB over PCIe (~20 ms) around 3.8 ms of
GPU work.  Attention is independent per (batch, KV-head group), so this entry point cuts the call into such chunks and
runs a three-stage pipeline on three CUDA streams with double-buffered device staging:

    copy-in stream : H2D q,k,v of chunk i+1      (one memcpy per tensor: a head range of an HND tensor is contiguous)
    compute stream : sageattn(chunk i)           (the caller's current stream; same kernels, same results bit for bit)
    copy-out stream: D2H o of chunk i-1          (PCIe is full duplex: overlaps the H2D of later chunks)

so the call costs about max(H2D time, D2H time, GPU time) plus one chunk of fill/drain instead of their sum.
Every chunk runs the unmodified `sageattn` path on its heads; per-head statistics (K mean, V scales, Q/K block scales)
do not depend on other heads, so the result equals the single-shot call exactly.
"""
from typing import Any, Optional
import torch

from .core import sageattn

_streams = {}


def _side_streams(dev: torch.device):
    key = (dev.type, dev.index)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _streams[key]


def _chunks(B: int, H: int, g: int, heads_per_chunk: int):
    out = []
    for b in range(B):
        for h0 in range(0, H, heads_per_chunk):
            out.append((b, h0, min(H, h0 + heads_per_chunk)))
    return out


def sageattn_host(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: Optional[torch.Tensor] = None,
                  tensor_layout: str = "HND", is_causal: bool = False, sm_scale: Optional[float] = None,
                  device: Optional[torch.device] = None, heads_per_chunk: Optional[int] = None, sync: bool = True,
                  **kwargs: Any) -> torch.Tensor:
    """q, k, v: CPU tensors (pinned memory gives asynchronous copies) in the layouts `sageattn` accepts
    (sageattention/core.py:79-88).  Returns `out` (a pinned CPU tensor shaped like q; allocated if not given).
    With sync=True (default) the result is complete on return; with sync=False the caller's current CUDA stream has
    been made to wait for the last D2H copy, so `torch.cuda.current_stream().synchronize()` completes it."""
    assert q.device.type == "cpu" and k.device.type == "cpu" and v.device.type == "cpu", "sageattn_host takes host tensors"
    assert q.dtype in (torch.float16, torch.bfloat16) and q.dtype == k.dtype == v.dtype, "q, k, v must be fp16 or bf16"
    if tensor_layout not in ("HND", "NHD"):
        raise ValueError(f"Unknown tensor layout: {tensor_layout}")
    if kwargs.get("return_lse", False):
        raise NotImplementedError("sageattn_host does not return the LSE")
    hnd = tensor_layout == "HND"
    B, H, S, D = (q.size(0), q.size(1), q.size(2), q.size(3)) if hnd else (q.size(0), q.size(2), q.size(1), q.size(3))
    Hkv = k.size(1) if hnd else k.size(2)
    Skv = k.size(2) if hnd else k.size(1)
    assert H % Hkv == 0, "num_qo_heads must be a multiple of num_kv_heads"
    g = H // Hkv
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if out is None:
        out = torch.empty(q.shape, dtype=q.dtype, pin_memory=True)
    assert out.device.type == "cpu" and out.shape == q.shape and out.dtype == q.dtype

    # ---- chunking: whole KV-head groups of one batch entry, ~48 MB of input per chunk.  NHD interleaves the heads of
    #      a token in memory, so an NHD call is cut along the batch only (a head range would be a strided copy).
    if not hnd:
        hc = H
    elif heads_per_chunk is not None:
        hc = max(g, (int(heads_per_chunk) // g) * g)
    else:
        per_group = (g * S + 2 * Skv) * D * q.element_size()       # q heads of one KV group + its k + its v
        hc = g * max(1, min(Hkv, (48 << 20) // max(per_group, 1)))
    chunks = _chunks(B, H, g, hc)

    with torch.cuda.device(dev):
        compute = torch.cuda.current_stream(dev)
        s_in, s_out = _side_streams(dev)
        nslot = 2
        qshape = (1, hc, S, D) if hnd else (1, S, H, D)
        kshape = (1, hc // g, Skv, D) if hnd else (1, Skv, Hkv, D)
        qd = [torch.empty(qshape, dtype=q.dtype, device=dev) for _ in range(nslot)]
        kd = [torch.empty(kshape, dtype=q.dtype, device=dev) for _ in range(nslot)]
        vd = [torch.empty(kshape, dtype=q.dtype, device=dev) for _ in range(nslot)]
        ev_in = [torch.cuda.Event() for _ in range(nslot)]      # H2D of the slot landed
        ev_free = [torch.cuda.Event() for _ in range(nslot)]    # compute has consumed the slot's inputs
        # the staging buffers were allocated on the compute stream; the copy-in stream must not run ahead of that
        s_in.wait_stream(compute)
        last_out = None
        for i, (b, h0, h1) in enumerate(chunks):
            slot = i % nslot
            nh = h1 - h0
            if hnd:
                qs, ks, vs = q[b:b + 1, h0:h1], k[b:b + 1, h0 // g:h1 // g], v[b:b + 1, h0 // g:h1 // g]
                qv, kv_, vv = qd[slot][:, :nh], kd[slot][:, :nh // g], vd[slot][:, :nh // g]
                ov = out[b:b + 1, h0:h1]
            else:
                qs, ks, vs = q[b:b + 1], k[b:b + 1], v[b:b + 1]
                qv, kv_, vv = qd[slot], kd[slot], vd[slot]
                ov = out[b:b + 1]
            with torch.cuda.stream(s_in):
                if i >= nslot:
                    s_in.wait_event(ev_free[slot])
                qv.copy_(qs, non_blocking=True)
                kv_.copy_(ks, non_blocking=True)
                vv.copy_(vs, non_blocking=True)
                ev_in[slot].record(s_in)
            compute.wait_event(ev_in[slot])
            o = sageattn(qv, kv_, vv, tensor_layout=tensor_layout, is_causal=is_causal, sm_scale=sm_scale, **kwargs)
            ev_free[slot].record(compute)
            ev_o = torch.cuda.Event()
            ev_o.record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_o)
                ov.copy_(o, non_blocking=True)
                o.record_stream(s_out)          # keep the allocator from reusing o before the D2H copy has read it
                last_out = torch.cuda.Event()
                last_out.record(s_out)
        # the staging buffers go back to the allocator on the compute stream: order that after their last H2D writes
        compute.wait_stream(s_in)
        if last_out is not None:
            compute.wait_event(last_out)
            if sync:
                last_out.synchronize()
    return out

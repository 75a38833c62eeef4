"""Kernel-only timing of the attention kernel on the headline shapes (quick A/B harness; SAB_LIB_PATH selects the .so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sageattention_b200 as sab
from sageattention_b200 import ops

dev = torch.device("cuda:0")
shapes = [(2, 32, 8192, 128, 0, "per_thread"), (2, 32, 8192, 128, 0, "per_warp"), (2, 32, 8192, 128, 1, "per_thread"), (2, 32, 8192, 64, 0, "per_thread"),
          (1, 32, 32768, 64, 1, "per_thread")]
if len(sys.argv) > 1 and sys.argv[1] == "short":
    shapes = shapes[:1]
if len(sys.argv) > 1 and sys.argv[1] == "short2":
    shapes = shapes[:3]
if len(sys.argv) > 1 and sys.argv[1] == "long":     # sequence-length sweep at head_dim 128 (kernel only)
    shapes = [(2, 32, 8192, 128, 0, "per_thread"), (1, 32, 16384, 128, 0, "per_thread"), (1, 32, 32768, 128, 0, "per_thread"),
              (1, 32, 16384, 128, 1, "per_thread"), (1, 32, 32768, 128, 1, "per_thread")]
out = []
for (B, H, S, D, causal, gran) in shapes:
    torch.manual_seed(0)
    q, k, v = (torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
    km = sab.k_mean(k)
    q8, qs, k8, ks = (sab.per_warp_int8 if gran == "per_warp" else sab.per_thread_int8)(q, k, km)
    v8, vs, _ = sab.per_channel_fp8(v, scale_max=2.25, smooth_v=False)
    o = torch.empty_like(q)
    g = 2 if gran == "per_warp" else 3
    f = lambda: ops.qk_int8_sv_f8_attn(q8, k8, v8, o, qs, ks, vs, None, 1, causal, g, g, D ** -0.5, 0, 0)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if S <= 8192 else 4
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * B * H * S * S * D / (2 if causal else 1)
    out.append(f"D{D} S{S} c{causal} {gran}: {fl / ms / 1e9:.0f}")
if not (len(sys.argv) > 1 and (sys.argv[1].startswith("short") or sys.argv[1] == "long")):
    # BASELINE configs[3]: sageattn_varlen, GQA Hq=32 / Hkv=8, hd=128, sequence lengths 512..16384 (whole call: quantisation + FP16-PV kernel)
    lens = [512, 1024, 2048, 4096, 8192, 16384]
    g = torch.Generator().manual_seed(0)
    lens = [lens[i] for i in torch.randperm(len(lens), generator=g).tolist()]
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    T = int(cu[-1])
    q = torch.randn(T, 32, 128, device=dev, dtype=torch.bfloat16)
    k = torch.randn(T, 8, 128, device=dev, dtype=torch.bfloat16)
    v = torch.randn(T, 8, 128, device=dev, dtype=torch.bfloat16)
    for causal in (False, True):
        f = lambda: sab.sageattn_varlen(q, k, v, cu, cu, max(lens), max(lens), is_causal=causal)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = sum(4.0 * 32 * L * L * 128 for L in lens) / (2 if causal else 1)
        out.append(f"varlen GQA c{int(causal)} (whole call): {fl / ms / 1e9:.0f}")
print(os.environ.get("SAB_LIB_PATH", "default"), " | ".join(out), flush=True)

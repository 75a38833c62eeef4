"""Front-end timing on the GPU box: the fused single-pass K / V kernels against the two-step entry points (configs[1] shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sageattention_b200 as sab
from sageattention_b200 import ops
from sageattention_b200.quant import quant_k_int8, quant_q_int8, smooth_quant_k

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for (B, H, S, D) in [(4, 32, 8192, 128), (1, 32, 32768, 128), (4, 32, 8192, 64)]:
    q, k, v = (torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    nb = B * H * S * D
    km = sab.k_mean(k)
    v8 = torch.empty((B, H, D, S), dtype=torch.float8_e4m3fn, device="cuda"); vs = torch.empty((B, H, D), dtype=torch.float32, device="cuda")
    r = {
        "quant Q": t(lambda: quant_q_int8(q, "per_thread")),
        "k_mean": t(lambda: sab.k_mean(k)),
        "quant K (given mean)": t(lambda: quant_k_int8(k, km, "per_thread")),
        "fused K (mean+quant)": t(lambda: smooth_quant_k(k, "per_thread")),
        "V two-step": t(lambda: ops.per_channel_fp8(v, v8, vs, None, 1, 2.25)),
        "V fused": t(lambda: ops.per_channel_fp8_fused(v, v8, vs, None, 1, 2.25)),
    }
    print(f"B{B} H{H} S{S} D{D} (SAB_FUSED_DYNSMEM={os.environ.get('SAB_FUSED_DYNSMEM', 'default')}): " +
          " | ".join(f"{k_}: {v_:.1f} us ({3 * nb / v_ / 1e6:.2f} TB/s of 3 B/elt)" for k_, v_ in r.items()), flush=True)

"""Launch the attention kernel a few times on a fixed shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sageattention_b200 as sab
from sageattention_b200 import ops

B, H, S, D = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1, 32, 8192, 128))]
causal = int(sys.argv[5]) if len(sys.argv) > 5 else 0
gran = sys.argv[6] if len(sys.argv) > 6 else "per_thread"
n = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
q = torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16)
k = torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16)
v = torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16)
km = sab.k_mean(k)
q8, qs, k8, ks = (sab.per_warp_int8 if gran == "per_warp" else sab.per_thread_int8)(q, k, km)
v8, vs, _ = sab.per_channel_fp8(v, scale_max=2.25, smooth_v=False)
o = torch.empty_like(q)
g = 2 if gran == "per_warp" else 3
for _ in range(n):
    ops.qk_int8_sv_f8_attn(q8, k8, v8, o, qs, ks, vs, None, 1, causal, g, g, D ** -0.5, 0, 0)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))

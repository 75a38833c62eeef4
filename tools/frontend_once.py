"""One pass of the quantisation front-end at configs[1] (for ncu: every kernel launched a few times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sageattention_b200 as sab
from sageattention_b200 import ops
from sageattention_b200.quant import quant_k_int8, quant_q_int8

B, H, S, D = 4, 32, 8192, 128
q, k, v = (torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
v8 = torch.empty((B, H, D, S), dtype=torch.float8_e4m3fn, device="cuda"); vs = torch.empty((B, H, D), dtype=torch.float32, device="cuda")
for _ in range(2):
    km = sab.k_mean(k)
    quant_q_int8(q, "per_thread")
    quant_k_int8(k, km, "per_thread")
    ops.per_channel_fp8(v, v8, vs, None, 1, 2.25)
torch.cuda.synchronize()

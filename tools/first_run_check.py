"""First-run check of a kernel build on the GPU box (run under `timeout`; SAB_LIB_PATH selects the library): the INT8+FP8 path on
small shapes against the CPU oracle evaluated with the kernel's own arithmetic (lazy max, tau = 3) and with the exact max."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sageattention_b200 as sab
from oracle import sage_oracle as O

tau = None if os.environ.get("SAB_ATTN_KERNEL", "")[:1] == "e" else 4
worst = 0.0
for (B, H, Hk, S, D, dt, causal, gran) in [(1, 2, 2, 64, 128, torch.float16, False, "per_thread"), (1, 2, 2, 128, 128, torch.float16, True, "per_warp"),
                                            (1, 2, 1, 320, 128, torch.bfloat16, False, "per_thread"), (2, 4, 2, 200, 128, torch.bfloat16, True, "per_thread"),
                                            (1, 2, 2, 1000, 128, torch.float16, True, "per_warp"), (1, 2, 2, 2048, 128, torch.float16, False, "per_thread"),
                                            (1, 1, 1, 1, 128, torch.float16, False, "per_thread"), (1, 2, 2, 77, 72, torch.float16, False, "per_thread")]:
    g = torch.Generator(device="cuda").manual_seed(S)
    q = torch.randn(B, H, S, D, device="cuda", generator=g).to(dt)
    k = (torch.randn(B, Hk, S, D, device="cuda", generator=g) + 4.0 * torch.randn(B, Hk, 1, D, device="cuda", generator=g)).to(dt)
    v = torch.randn(B, Hk, S, D, device="cuda", generator=g).to(dt)
    o, lse = sab.sageattn_qk_int8_pv_fp8_cuda(q, k, v, is_causal=causal, qk_quant_gran=gran, return_lse=True)
    torch.cuda.synchronize()
    kw = dict(is_causal=causal, qk_quant_gran=gran, return_lse=True, emulate_f16_accum=False)
    oe, le = O.sageattn_qk_int8_pv_fp8_cuda(q.cpu(), k.cpu(), v.cpu(), lazy_tau=tau, **kw)
    ox, lx = O.sageattn_qk_int8_pv_fp8_cuda(q.cpu(), k.cpu(), v.cpu(), **kw)
    sd = O.sdpa_fp32(q.cpu(), k.cpu(), v.cpu(), is_causal=causal)
    e1 = (o.cpu().float() - oe.float()).abs().max().item()
    e2 = (o.cpu().float() - ox.float()).abs().max().item()
    e3 = (o.cpu().float() - sd).abs().max().item()
    e3r = (ox.float() - sd).abs().max().item()
    el = (lse.cpu() - le).abs().max().item()
    worst = max(worst, e1)
    print(f"B{B} H{H}/{Hk} S{S} D{D} {str(dt)[6:]} c{int(causal)} {gran}: vs oracle(same arithmetic) {e1:.2e}  vs oracle(exact max) {e2:.2e}  "
          f"vs fp32 SDPA {e3:.2e} (exact-max oracle: {e3r:.2e})  lse {el:.2e}  nan={bool(torch.isnan(o).any())}", flush=True)
print("WORST", worst, flush=True)

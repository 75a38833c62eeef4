"""Mismatch report: our CUDA-semantics quantisers vs the real reference kernels (oracle/_ref/ref_fused.so)."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sageattention_b200 as sab

p = os.path.join(ROOT, "oracle", "_ref", "ref_fused.so")
spec = importlib.util.spec_from_file_location("ref_fused", p)
rf = importlib.util.module_from_spec(spec); spec.loader.exec_module(rf)
for (B, H, S, D, dt) in [(2, 3, 333, 128, torch.bfloat16), (1, 2, 1000, 64, torch.float16), (1, 2, 1024, 128, torch.float16)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(B, H, S, D, device="cuda", dtype=dt) for _ in range(3))
    k = k + torch.randn(1, H, 1, D, device="cuda", dtype=dt) * 3
    km = k.mean(dim=2, keepdim=True)
    q8, qs, k8, ks = sab.per_warp_int8(q, k, km)
    rq8, rk8, rqs, rks = torch.empty_like(q8), torch.empty_like(k8), torch.empty_like(qs), torch.empty_like(ks)
    rf.quant_per_warp_int8_cuda(q, rq8, rqs, 128, 32, 1)
    rf.quant_per_block_int8_fuse_sub_mean_cuda(k, km.squeeze(2), rk8, rks, 64, 1)
    torch.cuda.synchronize()
    for name, a, b_ in (("q8", q8, rq8), ("qs", qs, rqs), ("k8", k8, rk8), ("ks", ks, rks)):
        ne = (a != b_)
        n = int(ne.sum())
        msg = f"{(B, H, S, D, dt)} {name}: {n} / {a.numel()} differ"
        if n:
            idx = ne.nonzero()[:5].tolist()
            msg += f" first {idx} ours {[a[tuple(i)].item() for i in idx]} ref {[b_[tuple(i)].item() for i in idx]}"
        print(msg, flush=True)

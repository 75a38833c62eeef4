"""In-kernel clock64 timeline of the lazy kernel (lib built with -DSAB_TIMELINE) for tiles 16..47 of CTA (0,0,0), softmax warp 0 / MMA warp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sageattention_b200 as sab
from sageattention_b200 import _capi
dev = torch.device("cuda:0")
B, H, S, D = 2, 32, 8192, 128
torch.manual_seed(0)
q, k, v = (torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
km = sab.k_mean(k)
q8, qs, k8, ks = sab.per_thread_int8(q, k, km)
v8, vs, _ = sab.per_channel_fp8(v, scale_max=2.25, smooth_v=False)
o = torch.empty_like(q)
dbg = torch.zeros(128 * 64 + 128 * 16 + 128 * D + 256 + 8192 + 4096, dtype=torch.int32, device=dev)
for it in range(3):
    st = _capi.lib().sab_qk_int8_sv_f8_attn(q8.data_ptr(), k8.data_ptr(), v8.data_ptr(), o.data_ptr(), None, qs.data_ptr(), ks.data_ptr(), vs.data_ptr(), None, 1,
        B, H, H, S, S, D, q8.stride(0), q8.stride(1), q8.stride(2), k8.stride(0), k8.stride(1), k8.stride(2), v8.size(-1), o.stride(0), o.stride(1), o.stride(2),
        0, 3, 3, D ** -0.5, 0, None, None, None, None, None, 0, 0, 0, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _capi.check(st)
torch.cuda.synchronize()
tl = dbg.view(torch.int64)[4096:4096 + 32 * 16].view(32, 16).cpu()
t0 = tl[0, 0].item()
print("slots: 0 tile top | 5 prefetch: before wait | 6 after wait | 1 exp loop done | 2 committed | 3 wait::st done | 4 arrived ; mma 8 top 9 P ok 10 PV issued 11 QK issued+commit")
for j in range(0, 24):
    r = tl[j]
    nxt = int(tl[j + 1, 0] - r[0]) if j + 1 < 32 else -1
    print(f"tile {16 + j}: start {int(r[0] - t0):7d} | to-prefetch {int(r[5] - r[0]):5d} wait {int(r[6] - r[5]):5d} rest-of-exp {int(r[1] - r[6]):5d} vote {int(r[2] - r[1]):5d} "
          f"wait::st {int(r[3] - r[2]):5d} arrive {int(r[4] - r[3]):5d} | tile {nxt:6d} || mma: wait_p {int(r[9] - r[8]):6d} pv {int(r[10] - r[9]):5d} qk+commit {int(r[11] - r[10]):5d} "
          f"| P arrive->mma sees it {int(r[9] - r[4]):6d}")

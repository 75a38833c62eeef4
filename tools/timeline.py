"""Dump the in-kernel clock64 timeline (lib built with -DSAB_TIMELINE) for tiles 16..47 of CTA (0,0,0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sageattention_b200 as sab
from sageattention_b200 import _capi
dev = torch.device("cuda:0")
B, H, S, D = 2, 32, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
q, k, v = (torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
km = sab.k_mean(k)
q8, qs, k8, ks = sab.per_thread_int8(q, k, km)
v8, vs, _ = sab.per_channel_fp8(v, scale_max=2.25, smooth_v=False)
o = torch.empty_like(q)
dbg = torch.zeros(4096 * 2 + 2 * 32 * 16 * 2 + 1024, dtype=torch.int32, device=dev)
# dbg region [0, 128*64+128*16+128*D+256) is used by the debug dump: enlarge
dbg = torch.zeros(128 * 64 + 128 * 16 + 128 * D + 256 + 8192 + 4096, dtype=torch.int32, device=dev)
for it in range(3):
    st = _capi.lib().sab_qk_int8_sv_f8_attn(q8.data_ptr(), k8.data_ptr(), v8.data_ptr(), o.data_ptr(), None, qs.data_ptr(), ks.data_ptr(), vs.data_ptr(), None, 1,
        B, H, H, S, S, D, q8.stride(0), q8.stride(1), q8.stride(2), k8.stride(0), k8.stride(1), k8.stride(2), v8.size(-1), o.stride(0), o.stride(1), o.stride(2),
        0, 3, 3, D ** -0.5, 0, None, None, None, None, None, 0, 0, 0, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _capi.check(st)
torch.cuda.synchronize()
tl = dbg.view(torch.int64)[4096:4096 + 32 * 16].view(32, 16).cpu()
names = ["top", "s_full ok", "S loaded", "max/alpha", "exp done", "P st issued(+resc)", "wait::st", "arrived", "mma:top", "mma:P ok", "mma:PV issued", "mma:QK issued"]
t0 = tl[0, 0].item()
print("softmax warp0 lane0 timeline, cycles relative; per tile deltas")
for j in range(0, 20):
    r = tl[j]
    d = [int(r[i + 1] - r[i]) for i in range(7)]
    nxt = int(tl[j + 1, 0] - r[0]) if j + 1 < 32 else -1
    m = [int(r[9] - r[8]), int(r[10] - r[9]), int(r[11] - r[10])]
    st = [int(r[13] - r[12]), int(r[14] - r[13]), int(r[15] - r[14])]
    print(f"tile {16 + j}: start {int(r[0] - t0):7d} | wait_m {d[0]:5d} ldS {d[1]:5d} alpha {d[2]:5d} exp {d[3]:5d} st {d[4]:5d} waitst {d[5]:5d} arrive {d[6]:5d} | iter {nxt:6d} || mma wait_p {m[0]:6d} pv {m[1]:5d} qk {m[2]:5d} P_ok-arrive {int(r[9] - r[7]):6d} || stats wait_s {st[0]:5d} max+pub {st[1]:5d} resc+arrive {st[2]:5d} m_pub-exp_top {int(r[14] - r[0]):6d}")

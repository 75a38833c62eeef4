"""a16 (two-level f16/f32 PV accumulation, csrc/qattn/attn_utils.cuh:896-974) evidence, run on the GPU box:
  (i)  the oracle's emulate_f16_accum branch against the REAL reference "fp32+fp16" kernel (oracle/_ref/ref_qattn.so);
  (ii) on the rows where the B200 kernel (fp32 accumulation in TMEM) differs most from that kernel, which of the two is
       closer to the exact evaluation of the same quantised operands (oracle, fp32 accumulation) and to fp32 SDPA.
Prints one line per case; tests/test_gpu_parity.py asserts the same quantities."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sageattention_b200 as sab
from sageattention_b200 import ops
from oracle import sage_oracle as O


def _ref(name):
    p = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    spec = importlib.util.spec_from_file_location(name, p)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m


rf, ra = _ref("ref_fused"), _ref("ref_qattn")
CASES = [(1, 4, 1024, 128, torch.float16, False, "per_warp", 0.0), (1, 4, 1024, 64, torch.bfloat16, True, "per_warp", 0.0),
         (1, 2, 2048, 128, torch.bfloat16, True, "per_thread", 0.0), (1, 2, 333, 128, torch.float16, True, "per_thread", 0.0),
         (1, 2, 2048, 128, torch.float16, False, "per_thread", 2.0), (1, 2, 1024, 128, torch.float16, True, "per_thread", 2.0)]
for (B, H, S, D, dt, causal, gran, vshift) in CASES:
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, H, S, D, device="cuda", generator=g).to(dt)
    k = (torch.randn(B, H, S, D, device="cuda", generator=g) + 4.0 * torch.randn(B, H, 1, D, device="cuda", generator=g)).to(dt)
    v = (torch.randn(B, H, S, D, device="cuda", generator=g) + vshift).to(dt)     # v += 2: the f16-accumulator stress of SURVEY §8(d)
    km = k.mean(dim=2, keepdim=True)
    sm = D ** -0.5
    q8, qs, k8, ks = (sab.per_warp_int8 if gran == "per_warp" else sab.per_thread_int8)(q, k, km)
    gi = 2 if gran == "per_warp" else 3
    pl = (S + 63) // 64 * 64
    vt = torch.empty((B, H, D, pl), dtype=dt, device="cuda")
    rf.transpose_pad_permute_cuda(v, vt, 1)
    r8 = torch.empty(vt.shape, dtype=torch.float8_e4m3fn, device="cuda"); rs = torch.empty((B, H, D), dtype=torch.float32, device="cuda")
    rf.scale_fuse_quant_cuda(vt, r8, rs, S, 2.25, 1)
    o_ref = torch.empty_like(q)
    ra.qk_int8_sv_f8_accum_f16_fuse_v_scale_attn_inst_buf(q8, k8, r8, o_ref, qs, ks, rs, 1, int(causal), gi, sm, 0)
    v8, vs, _ = sab.per_channel_fp8(v, scale_max=2.25, smooth_v=False)
    o = torch.empty_like(q)
    ops.qk_int8_sv_f8_attn(q8, k8, v8, o, qs, ks, vs, None, 1, int(causal), gi, gi, sm, 0, 0)
    torch.cuda.synchronize()
    # the same quantised operands on the CPU: v8 logical [B,H,S,D]
    v8l = v8[..., :S].transpose(2, 3).contiguous().cpu()
    args = (q8.cpu(), k8.cpu(), v8l, qs.cpu(), ks.cpu(), vs.cpu())
    kw = dict(qk_quant_gran=gran, is_causal=causal, sm_scale=sm, out_dtype=torch.float32)
    o_f16 = O.attn_int8_fp8_cuda(*args, pv_accum_dtype="fp32+fp16", **kw)
    o_f32 = O.attn_int8_fp8_cuda(*args, pv_accum_dtype="fp32+fp32", **kw)
    sd = O.sdpa_fp32(q.cpu(), k.cpu(), v.cpu(), is_causal=causal)
    of, rf_, = o.float().cpu(), o_ref.float().cpu()
    diff = (of - rf_).abs()
    rows = diff.amax(dim=-1) > 5e-3
    e_o_exact, e_r_exact = (of - o_f32).abs(), (rf_ - o_f32).abs()
    e_o_sd, e_r_sd = (of - sd).abs(), (rf_ - sd).abs()
    print(f"B{B} H{H} S{S} D{D} {str(dt)[6:]} causal={int(causal)} {gran} v+{vshift}: |ours-ref16| max {diff.max():.2e}; "
          f"|oracle_f16 - ref16| max {(o_f16 - rf_).abs().max():.2e}; |oracle_f32 - ours| max {(o_f32 - of).abs().max():.2e}; "
          f"rows>5e-3: {int(rows.sum())}; on them vs exact-quantised: ours {e_o_exact[rows].max() if rows.any() else 0:.2e} "
          f"ref {e_r_exact[rows].max() if rows.any() else 0:.2e}; all rows vs exact-quantised: ours max {e_o_exact.max():.2e} mean {e_o_exact.mean():.2e}, "
          f"ref max {e_r_exact.max():.2e} mean {e_r_exact.mean():.2e}; vs fp32 SDPA: ours max {e_o_sd.max():.2e} mean {e_o_sd.mean():.2e}, "
          f"ref max {e_r_sd.max():.2e} mean {e_r_sd.mean():.2e}", flush=True)

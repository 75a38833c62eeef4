"""GPU bring-up probe (run under gpurun): exercises every kernel against the CPU oracle and, when
oracle/_ref/*.so is present, against the real reference CUDA kernels.  Prints diagnostics instead of
stopping at the first failure.  Not part of the product."""
import os, sys, time, importlib.util, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sageattention_b200 as sab
from sageattention_b200 import ops, _capi
from sageattention_b200.quant import per_block_int8_varlen, per_channel_fp8_varlen
from oracle import sage_oracle as O

dev = torch.device("cuda:0")
torch.manual_seed(0)


def load_ref(name):
    p = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location(name, p)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m


def section(t):
    print("\n" + "=" * 20, t, "=" * 20, flush=True)


def guarded(fn):
    def w(*a, **k):
        try:
            return fn(*a, **k)
        except Exception:
            traceback.print_exc()
            torch.cuda.synchronize()
    return w


def mk(B, H, S, D, dtype, outlier=True, Hk=None):
    Hk = Hk or H
    q = torch.randn(B, H, S, D, device=dev).to(dtype)
    k = torch.randn(B, Hk, S, D, device=dev)
    if outlier:
        k = k + 4.0 * torch.randn(B, Hk, 1, D, device=dev)
    v = torch.randn(B, Hk, S, D, device=dev)
    return q, k.to(dtype), v.to(dtype)


@guarded
def probe_quant():
    section("quant vs oracle")
    for (B, H, S, D, dt) in [(1, 2, 200, 64, torch.float16), (2, 3, 333, 128, torch.bfloat16), (1, 2, 1024, 128, torch.float16)]:
        q, k, v = mk(B, H, S, D, dt)
        km = sab.k_mean(k)
        km_ref = k.mean(dim=2, keepdim=True)
        print(f"[{B},{H},{S},{D},{dt}] k_mean mismatches vs torch: {(km != km_ref).sum().item()} / {km.numel()}  maxdiff {(km.float()-km_ref.float()).abs().max().item():.3e}")
        qc, kc, kmc = q.cpu(), k.cpu(), km_ref.cpu()
        # triton-semantics per-block + per-thread: bit exact expected
        for name, fn, ofn in [("per_block(triton)", lambda: sab.per_block_int8(q, k, km_ref, sm_scale=D ** -0.5),
                               lambda: O.per_block_int8_triton(qc, kc, kmc, sm_scale=D ** -0.5)),
                              ("per_thread", lambda: sab.per_thread_int8(q, k, km_ref), lambda: O.quant_per_thread_int8_triton(qc, kc, kmc)),
                              ("per_warp(cuda)", lambda: sab.per_warp_int8(q, k, km_ref), lambda: O.per_warp_int8_cuda(qc, kc, kmc))]:
            got = [t.cpu() for t in fn()]
            exp = ofn()
            msg = []
            for nm, g, e in zip(["q8", "qs", "k8", "ks"], got, exp):
                if g.dtype == torch.int8:
                    diff = (g.int() - e.int()).abs()
                    msg.append(f"{nm}: mism {int((diff > 0).sum())} max {int(diff.max())}")
                else:
                    msg.append(f"{nm}: mism {int((g != e).sum())}/{g.numel()} maxrel {((g - e).abs() / e.abs().clamp_min(1e-20)).max().item():.2e}")
            print("   ", name, " | ".join(msg))
        # V
        for smax in (448.0, 2.25):
            v8, vs, _ = sab.per_channel_fp8(v, scale_max=smax, smooth_v=False)
            e8, es, _ = O.per_channel_fp8_cuda(v.cpu(), "HND", smax)
            got = v8[:, :, :, :S].transpose(2, 3).float().cpu()
            pad = v8[:, :, :, S:].float().abs().sum().item()
            d = (got - e8.float()).abs()
            print(f"    V fp8 smax={smax}: mism {int((d > 0).sum())}/{d.numel()} maxabs {d.max().item():.3e} pad_sum {pad}  scale mism {int((vs.cpu() != es).sum())} maxrel {((vs.cpu()-es).abs()/es).max().item():.2e}")


@guarded
def probe_ref_quant():
    section("quant vs REAL reference kernels (oracle/_ref)")
    rf = load_ref("ref_fused")
    if rf is None:
        print("ref_fused.so absent"); return
    for (B, H, S, D, dt) in [(2, 3, 333, 128, torch.bfloat16), (1, 2, 1000, 64, torch.float16)]:
        q, k, v = mk(B, H, S, D, dt)
        km = k.mean(dim=2, keepdim=True)
        q8, qs, k8, ks = sab.per_warp_int8(q, k, km)
        rq8 = torch.empty_like(q8); rk8 = torch.empty_like(k8); rqs = torch.empty_like(qs); rks = torch.empty_like(ks)
        rf.quant_per_warp_int8_cuda(q, rq8, rqs, 128, 32, 1)
        rf.quant_per_block_int8_fuse_sub_mean_cuda(k, km.squeeze(2), rk8, rks, 64, 1)
        torch.cuda.synchronize()
        print(f"[{B},{H},{S},{D},{dt}] per_warp vs ref: q8 mism {int((q8 != rq8).sum())} qs mism {int((qs != rqs).sum())} k8 mism {int((k8 != rk8).sum())} ks mism {int((ks != rks).sum())}")
        if int((qs != rqs).sum()):
            i = (qs != rqs).nonzero()[0]
            print("      first qs diff", qs[tuple(i)].item(), rqs[tuple(i)].item())
        # per-block with sm_scale (cuda semantics)
        q8b, qsb, k8b, ksb = sab.per_block_int8(q, k, km, sm_scale=D ** -0.5, semantics="cuda")
        rq8 = torch.empty_like(q8b); rqs = torch.empty_like(qsb)
        rf.quant_per_block_int8_scale_cuda(q, rq8, rqs, (D ** -0.5) * 1.44269504, 128, 1)
        torch.cuda.synchronize()
        print(f"      per_block(cuda, sm_scale) vs ref: q8 mism {int((q8b != rq8).sum())} qs mism {int((qsb != rqs).sum())}")
        # V: reference layout = transposed, padded 64, permuted within 16
        for smax in (448.0, 2.25):
            v8, vs, _ = sab.per_channel_fp8(v, scale_max=smax, smooth_v=False)
            pl = (S + 63) // 64 * 64
            vt = torch.empty((B, H, D, pl), dtype=dt, device=dev)
            rf.transpose_pad_permute_cuda(v, vt, 1)
            r8 = torch.empty(vt.shape, dtype=torch.float8_e4m3fn, device=dev); rs = torch.empty_like(vs)
            rf.scale_fuse_quant_cuda(vt, r8, rs, S, smax, 1)
            torch.cuda.synchronize()
            perm = torch.tensor([0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15], device=dev)
            idx = (torch.arange(pl, device=dev) // 16) * 16
            # position p holds token perm[p%16] of its 16-group
            src = idx + perm[torch.arange(pl, device=dev) % 16]
            unperm = torch.empty(pl, dtype=torch.long, device=dev); unperm[src] = torch.arange(pl, device=dev)
            r_log = r8.view(torch.uint8)[..., unperm][..., :S]
            mine = v8.view(torch.uint8)[..., :S]
            print(f"      V fp8 smax={smax} vs ref: bytes mism {int((r_log != mine).sum())}/{mine.numel()}  scale mism {int((vs != rs).sum())}")


def call_attn_debug(q8, k8, v8, qs, ks, vs, gran, causal, sm_scale, dt):
    B, H, S, D = q8.shape
    Hk = k8.shape[1]
    o = torch.empty((B, H, S, D), dtype=dt, device=dev)
    dbg = torch.zeros(128 * 64 + 128 * 16 + 128 * D + 256, dtype=torch.int32, device=dev)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=dev)
    st = _capi.lib().sab_qk_int8_sv_f8_attn(q8.data_ptr(), k8.data_ptr(), v8.data_ptr(), o.data_ptr(), lse.data_ptr(),
        qs.data_ptr(), ks.data_ptr(), vs.data_ptr(), None, 0 if dt == torch.float16 else 1, B, H, Hk, S, k8.shape[2], D,
        q8.stride(0), q8.stride(1), q8.stride(2), k8.stride(0), k8.stride(1), k8.stride(2), v8.size(-1),
        o.stride(0), o.stride(1), o.stride(2), causal, gran, gran, sm_scale, 0, None, None, None, None, None, 0, 0, 0,
        dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _capi.check(st)
    torch.cuda.synchronize()
    return o, lse, dbg


@guarded
def probe_attn_debug():
    section("attention kernel internals (debug dump of tile 0)")
    for D in (128, 64):
        B, H, S = 1, 1, 256
        dt = torch.float16
        q, k, v = mk(B, H, S, D, dt, outlier=False)
        q8, qs, k8, ks = sab.per_warp_int8(q, k, None)
        v8, vs, _ = sab.per_channel_fp8(v, scale_max=448.0, smooth_v=False)
        sm = D ** -0.5
        o, lse, dbg = call_attn_debug(q8, k8, v8, qs, ks, vs, 2, 0, sm, dt)
        S_got = dbg[:128 * 64].view(128, 64).cpu()
        S_exp = (q8[0, 0, :128].float() @ k8[0, 0, :64].float().T).int().cpu()
        bad = (S_got != S_exp)
        print(f"D={D}: QK^T int32 tile0 mismatches {int(bad.sum())}/8192", "" if not bad.any() else f"first bad {bad.nonzero()[0].tolist()} got {S_got[tuple(bad.nonzero()[0])].item()} exp {S_exp[tuple(bad.nonzero()[0])].item()}")
        if bad.any():
            print("   got row0[:16]", S_got[0, :16].tolist()); print("   exp row0[:16]", S_exp[0, :16].tolist())
            print("   rows bad", bad.any(1).sum().item(), "cols bad", bad.any(0).sum().item())
        qs_row = O._expand_q_scale(qs.cpu(), "per_warp", S)[0, 0, :128]
        ks_key = O._expand_k_scale(ks.cpu(), "per_warp", S)[0, 0, :64]
        X = S_exp.float() * (qs_row[:, None] * ks_key[None, :] * (sm * O.LOG2E_CU))
        m0 = X.amax(1) - 8.807
        P = torch.exp2(X - m0[:, None]).to(torch.float8_e4m3fn)
        P_got = dbg[128 * 64:128 * 64 + 128 * 16].view(128, 16).cpu().contiguous().view(torch.uint8).view(128, 64).view(torch.float8_e4m3fn)
        dP = (P_got.float() - P.float()).abs()
        print(f"      P(e4m3) tile0: exact-byte mism {int((P_got.view(torch.uint8) != P.view(torch.uint8)).sum())}/8192 maxabs {dP.max().item():.3f} (1 ulp at 448 = 32)")
        out_exp = O.attn_int8_fp8_cuda(q8.cpu(), k8.cpu(), v8[..., :S].transpose(2, 3).contiguous().cpu(), qs.cpu(), ks.cpu(), vs.cpu(),
                                       qk_quant_gran="per_warp", sm_scale=sm, pv_accum_dtype="fp32+fp32", kv_tile=64, out_dtype=dt, return_lse=True)
        print(f"      O maxabs vs oracle(kv_tile=64): {(o.cpu().float() - out_exp[0].float()).abs().max().item():.4e}   lse maxabs {(lse.cpu() - out_exp[1]).abs().max().item():.3e}")
        Oraw = dbg[128 * 64 + 128 * 16:128 * 64 + 128 * 16 + 128 * D].view(torch.float32).view(128, D).cpu()
        dd = dbg[128 * 64 + 128 * 16 + 128 * D:][:128].view(torch.float32).cpu()
        print(f"      raw O finite: {bool(torch.isfinite(Oraw).all())}  d range [{dd.min().item():.3e},{dd.max().item():.3e}]")


@guarded
def probe_attn():
    section("attention vs oracle (end-to-end API)")
    cfgs = [
        dict(B=1, H=2, S=320, D=64, dt=torch.float16, causal=False, gran="per_warp", acc="fp32+fp32"),
        dict(B=1, H=2, S=320, D=64, dt=torch.float16, causal=True, gran="per_thread", acc="fp32+fp16"),
        dict(B=2, H=4, S=200, D=128, dt=torch.bfloat16, causal=False, gran="per_thread", acc="fp32+fp16", Hk=2),
        dict(B=1, H=2, S=1000, D=128, dt=torch.float16, causal=True, gran="per_warp", acc="fp32+fp16"),
        dict(B=1, H=2, S=1024, D=128, dt=torch.bfloat16, causal=False, gran="per_thread", acc="fp32+fp32"),
        dict(B=1, H=2, S=77, D=72, dt=torch.float16, causal=False, gran="per_thread", acc="fp32+fp16"),
    ]
    for c in cfgs:
        q, k, v = mk(c["B"], c["H"], c["S"], c["D"], c["dt"], Hk=c.get("Hk"))
        o, lse = sab.sageattn_qk_int8_pv_fp8_cuda(q, k, v, is_causal=c["causal"], qk_quant_gran=c["gran"], pv_accum_dtype=c["acc"], return_lse=True)
        torch.cuda.synchronize()
        oe, le = O.sageattn_qk_int8_pv_fp8_cuda(q.cpu(), k.cpu(), v.cpu(), is_causal=c["causal"], qk_quant_gran=c["gran"], pv_accum_dtype=c["acc"], return_lse=True)
        ex = O.sdpa_fp32(q.cpu(), k.cpu(), v.cpu(), is_causal=c["causal"])
        print({k_: (str(v_) if k_ == 'dt' else v_) for k_, v_ in c.items()},
              f"\n     maxabs vs oracle {(o.cpu().float() - oe.float()).abs().max().item():.4e} | vs exact {(o.cpu().float() - ex).abs().max().item():.4e} (oracle vs exact {(oe.float() - ex).abs().max().item():.4e}) | lse maxabs {(lse.cpu() - le).abs().max().item():.3e} | nan {int(torch.isnan(o).sum())}")
        # NHD layout
        qn, kn, vn = (t.transpose(1, 2).contiguous() for t in (q, k, v))
        on = sab.sageattn_qk_int8_pv_fp8_cuda(qn, kn, vn, tensor_layout="NHD", is_causal=c["causal"], qk_quant_gran=c["gran"], pv_accum_dtype=c["acc"])
        print(f"     NHD vs HND maxabs {(on.transpose(1, 2).float() - o.float()).abs().max().item():.3e}")


@guarded
def probe_ref_attn():
    section("attention vs REAL reference kernel (oracle/_ref)")
    rf, ra = load_ref("ref_fused"), load_ref("ref_qattn")
    if rf is None or ra is None:
        print("ref .so absent"); return
    for (B, H, S, D, dt, causal, gran) in [(1, 4, 1024, 128, torch.float16, False, "per_warp"), (1, 4, 1024, 64, torch.bfloat16, True, "per_warp"),
                                           (2, 4, 2000, 128, torch.bfloat16, False, "per_thread")]:
        q, k, v = mk(B, H, S, D, dt)
        km = k.mean(dim=2, keepdim=True)
        sm = D ** -0.5
        if gran == "per_warp":
            q8, qs, k8, ks = sab.per_warp_int8(q, k, km)
        else:
            q8, qs, k8, ks = sab.per_thread_int8(q, k, km)
        for acc, smax, fn in [("fp32+fp16", 2.25, ra.qk_int8_sv_f8_accum_f16_fuse_v_scale_attn_inst_buf), ("fp32+fp32", 448.0, ra.qk_int8_sv_f8_accum_f32_fuse_v_scale_attn_inst_buf)]:
            pl = (S + 63) // 64 * 64
            vt = torch.empty((B, H, D, pl), dtype=dt, device=dev)
            rf.transpose_pad_permute_cuda(v, vt, 1)
            r8 = torch.empty(vt.shape, dtype=torch.float8_e4m3fn, device=dev); rs = torch.empty((B, H, D), dtype=torch.float32, device=dev)
            rf.scale_fuse_quant_cuda(vt, r8, rs, S, smax, 1)
            o_ref = torch.empty_like(q)
            lse_ref = fn(q8, k8, r8, o_ref, qs, ks, rs, 1, int(causal), 2 if gran == "per_warp" else 3, sm, 1)
            v8, vs, _ = sab.per_channel_fp8(v, scale_max=smax, smooth_v=False)
            o = torch.empty_like(q)
            g = 2 if gran == "per_warp" else 3
            lse = ops.qk_int8_sv_f8_attn(q8, k8, v8, o, qs, ks, vs, None, 1, int(causal), g, g, sm, 0, 1)
            torch.cuda.synchronize()
            ex = O.sdpa_fp32(q.cpu(), (k - km).cpu(), v.cpu(), is_causal=causal)
            print(f"[{B},{H},{S},{D},{dt},causal={causal},{gran},{acc}] maxabs ours-vs-ref {(o.float() - o_ref.float()).abs().max().item():.4e} | ours-vs-exact {(o.cpu().float() - ex).abs().max().item():.4e} | ref-vs-exact {(o_ref.cpu().float() - ex).abs().max().item():.4e} | lse {(lse - lse_ref).abs().max().item():.3e}")


@guarded
def probe_varlen():
    section("varlen vs oracle")
    lens = [200, 130, 77, 512]
    Hq, Hk, D = 4, 2, 128
    T = sum(lens)
    q = torch.randn(T, Hq, D, device=dev).half(); k = (torch.randn(T, Hk, D, device=dev) + 2 * torch.randn(1, Hk, D, device=dev)).half(); v = torch.randn(T, Hk, D, device=dev).half()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    for causal in (False, True):
        o = sab.sageattn_varlen(q, k, v, cu, cu, max(lens), max(lens), is_causal=causal)
        torch.cuda.synchronize()
        oe = O.sageattn_varlen(q.cpu(), k.cpu(), v.cpu(), cu.cpu(), cu.cpu(), max(lens), max(lens), is_causal=causal)
        print(f"causal={causal}: maxabs vs oracle(fp16-PV reference semantics) {(o.cpu().float() - oe.float()).abs().max().item():.4e}  nan {int(torch.isnan(o).sum())}")
    km = k.mean(dim=0, keepdim=True)
    q8, qs, k8, ks, cuqs, cuks = per_block_int8_varlen(q, k, cu, cu, max(lens), max(lens), sm_scale=D ** -0.5, km=km)
    eq8, eqs, ecuqs = O.quant_per_block_int8_varlen_triton(q.cpu(), cu.cpu(), 128, (1.0 / D ** 0.5) * O.LOG2E_PY)
    ek8, eks, ecuks = O.quant_per_block_int8_varlen_triton((k - km).cpu(), cu.cpu(), 64, 1.0)
    n1, n2 = eqs.shape[0], eks.shape[0]
    print(f"varlen quant: q8 mism {int((q8.cpu() != eq8).sum())} k8 mism {int((k8.cpu() != ek8).sum())} qs mism {int((qs[:n1].cpu() != eqs).sum())} ks mism {int((ks[:n2].cpu() != eks).sum())}")


def bench_attn(B, H, S, D, causal, gran, dt=torch.bfloat16, iters=20):
    q, k, v = mk(B, H, S, D, dt)
    km = k.mean(dim=2, keepdim=True)
    q8, qs, k8, ks = (sab.per_warp_int8 if gran == "per_warp" else sab.per_thread_int8)(q, k, km)
    v8, vs, _ = sab.per_channel_fp8(v, scale_max=2.25, smooth_v=False)
    o = torch.empty_like(q)
    g = 2 if gran == "per_warp" else 3
    f = lambda: ops.qk_int8_sv_f8_attn(q8, k8, v8, o, qs, ks, vs, None, 1, int(causal), g, g, D ** -0.5, 0, 0)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * S * S * D / (2 if causal else 1)
    # e2e
    fe = lambda: sab.sageattn_qk_int8_pv_fp8_cuda(q, k, v, is_causal=causal, qk_quant_gran=gran)
    for _ in range(3): fe()
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fe()
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"B{B} H{H} S{S} D{D} causal={causal} {gran}: kernel {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOPS | e2e sageattn {ms2:.3f} ms = {fl / ms2 / 1e9:.1f} TFLOPS", flush=True)
    return q, k, v, km


@guarded
def probe_perf():
    section("first timings")
    bench_attn(2, 32, 8192, 128, False, "per_warp")
    bench_attn(2, 32, 8192, 128, False, "per_thread")
    bench_attn(4, 32, 8192, 128, False, "per_thread")
    bench_attn(2, 32, 8192, 128, True, "per_thread")
    bench_attn(2, 32, 8192, 64, False, "per_thread")
    bench_attn(1, 32, 32768, 64, True, "per_thread", iters=5)
    rf, ra = load_ref("ref_fused"), load_ref("ref_qattn")
    if ra is not None:
        B, H, S, D = 2, 32, 8192, 128
        q, k, v = mk(B, H, S, D, torch.bfloat16)
        km = k.mean(dim=2, keepdim=True)
        q8, qs, k8, ks = sab.per_warp_int8(q, k, km)
        vt = torch.empty((B, H, D, S), dtype=torch.bfloat16, device=dev)
        rf.transpose_pad_permute_cuda(v, vt, 1)
        r8 = torch.empty(vt.shape, dtype=torch.float8_e4m3fn, device=dev); rs = torch.empty((B, H, D), dtype=torch.float32, device=dev)
        rf.scale_fuse_quant_cuda(vt, r8, rs, S, 2.25, 1)
        o_ref = torch.empty_like(q)
        f = lambda: ra.qk_int8_sv_f8_accum_f16_fuse_v_scale_attn_inst_buf(q8, k8, r8, o_ref, qs, ks, rs, 1, 0, 2, D ** -0.5, 0)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"REFERENCE sm89 kernel on B200, same shape per_warp fp32+fp16: {ms:.3f} ms = {4.0 * B * H * S * S * D / ms / 1e9:.1f} TFLOPS")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    which = sys.argv[1:] or ["quant", "refquant", "dbg", "attn", "refattn", "varlen", "perf"]
    table = dict(quant=probe_quant, refquant=probe_ref_quant, dbg=probe_attn_debug, attn=probe_attn, refattn=probe_ref_attn,
                 varlen=probe_varlen, perf=probe_perf)
    for w in which:
        table[w]()
    print("\nPROBE DONE")

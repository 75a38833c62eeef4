"""One-shot GPU validation of the attn_mask kernel variant (sab_qk_int8_sv_f16_attn_masked): runs the guarded pytest cases
in-process and, for every fixture, prints where the CUDA output departs from the reference-Triton fixture (row / column /
tile statistics) so a failure can be debugged offline.  Usage on the GPU box:
    python tools/mask_validate.py > gpurun_out/mask_validate.log 2>&1
(first run on a B200: profiles/r01_mask_validate.log — all three pytest cases pass, max |o - reference Triton| 1.95e-3.)"""
import os, sys, traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def detail():
    import sageattention_b200 as sab
    G = os.path.join(ROOT, "tests", "golden")
    for name in ("attn_mask_bool_d64", "attn_mask_bias_d128"):
        try:
            z = np.load(f"{G}/{name}.npz")
            t = lambda a: torch.from_numpy(a.copy()).view(torch.float16)
            q, k, v, o_ref = (t(z[n]).cuda() for n in ("q", "k", "v", "o"))
            shape = tuple(int(x) for x in z["mask_shape"])
            mask = torch.from_numpy(z["mask"].copy()).view(shape) if str(z["kind"]) == "bool" else t(z["mask"]).view(shape)
            o, lse = sab.sageattn_qk_int8_pv_fp16_triton(q, k, v, attn_mask=mask.cuda(), return_lse=True)
            o0 = sab.sageattn_qk_int8_pv_fp16_triton(q, k, v)
            torch.cuda.synchronize()
            err = (o.float() - o_ref.float()).abs()
            lerr = np.abs(lse.cpu().numpy() - z["lse"])
            print(f"[{name}] q{tuple(q.shape)} mask{shape} kind={z['kind']}: max|o-ref|={err.max().item():.3e} "
                  f"mean={err.mean().item():.3e} max|lse-ref|={lerr.max():.3e} nan={int(torch.isnan(o).sum())} "
                  f"(unmasked call differs from ref by {(o0.float() - o_ref.float()).abs().max().item():.3e})")
            rows = err.amax(dim=-1)                       # [B,H,S]
            bad = (rows > 4e-3).nonzero()
            print(f"   rows over 4e-3: {bad.shape[0]} of {rows.numel()}; first: {bad[:8].tolist()}")
            if bad.shape[0]:
                b, h, s = bad[0].tolist()
                print("   o   :", o[b, h, s, :8].tolist())
                print("   ref :", o_ref[b, h, s, :8].tolist())
                print("   lse :", float(lse[b, h, s]), "ref", float(z["lse"][b, h, s]))
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    import pytest
    rc = pytest.main([os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-k", "attn_mask", "-q", "-rA", "--tb=short",
                      "-p", "no:cacheprovider"])
    print("pytest rc:", rc)
    detail()

"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares,
validates arguments before touching CUDA, and the Python API mirrors the reference signatures."""
import inspect, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sageattention_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "sageattn_b200.h")).read()
    declared = set(re.findall(r"\b(sab_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sageattn_b200.h but not exported"
    assert declared == set(_capi.EXPORTS), f"ctypes table out of sync: {declared ^ set(_capi.EXPORTS)}"
    assert lib.sab_version() >= 100


def test_argument_validation_without_gpu():
    from sageattention_b200 import _capi
    lib = _capi.lib()
    st = lib.sab_quant_per_block_int8(None, 0, None, None, None, 1, 1, 1, 128, 0, 0, 0, 0, 0, 0, 1, 128, 0, 0, 1.0, None)
    assert st == -1 and b"null" in lib.sab_last_error()
    # head_dim outside {64,128} (reference: DISPATCH_HEAD_DIM throws, csrc/dispatch_utils.h:23-34)
    st = lib.sab_qk_int8_sv_f8_attn(*([1] * 4), None, 1, 1, None, None, 0, 1, 2, 2, 128, 128, 96, *([0] * 10), 0, 2, 2, 1.0, 0,
                                    *([None] * 5), 0, 0, 0, None, None)
    assert st == -2 and b"head dim" in lib.sab_last_error()
    # GQA divisibility (reference: sm89_...inst_buf.cu:102-106)
    st = lib.sab_qk_int8_sv_f8_attn(*([1] * 4), None, 1, 1, None, None, 0, 1, 3, 2, 128, 128, 128, *([0] * 10), 0, 2, 2, 1.0, 0,
                                    *([None] * 5), 0, 0, 0, None, None)
    assert st == -1 and b"divisible" in lib.sab_last_error()
    with pytest.raises(ValueError):
        _capi.check(st)


def test_python_api_mirrors_reference_signatures():
    import sageattention_b200 as sab
    # sageattention/__init__.py:1-5
    for name in ["sageattn", "sageattn_varlen", "sageattn_qk_int8_pv_fp16_triton", "sageattn_qk_int8_pv_fp16_cuda",
                 "sageattn_qk_int8_pv_fp8_cuda", "sageattn_qk_int8_pv_fp8_cuda_sm90"]:
        assert callable(getattr(sab, name))
    sig = inspect.signature(sab.sageattn)                       # core.py:79-88
    assert list(sig.parameters)[:7] == ["q", "k", "v", "tensor_layout", "is_causal", "sm_scale", "return_lse"]
    assert sig.parameters["tensor_layout"].default == "HND" and sig.parameters["is_causal"].default is False
    assert any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values())
    sig = inspect.signature(sab.sageattn_qk_int8_pv_fp8_cuda)   # core.py:636-649
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["qk_quant_gran"] == "per_thread" and d["pv_accum_dtype"] == "fp32+fp16" and d["smooth_k"] is True and d["smooth_v"] is False
    sig = inspect.signature(sab.sageattn_varlen)                # core.py:334-346
    assert list(sig.parameters)[:10] == ["q", "k", "v", "cu_seqlens_q", "cu_seqlens_k", "max_seqlen_q", "max_seqlen_k",
                                         "is_causal", "sm_scale", "smooth_k"]


def test_cpu_tensors_are_rejected_not_silently_computed():
    """There is no CPU fallback: the ops exist only for CUDA tensors (reference: `assert q.is_cuda`, core.py:725)."""
    import torch
    import sageattention_b200 as sab
    q = torch.randn(1, 2, 128, 64, dtype=torch.float16)
    with pytest.raises(AssertionError):
        sab.sageattn(q, q, q)
    with pytest.raises(Exception):
        sab.per_warp_int8(q, q)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "sageattention_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/_ref", "").replace("the oracle", "").lower() or f == "build.py", f


def test_extensions_beyond_the_reference_surface_are_exported_and_guarded():
    """quantize_kv / sageattn_prequantized / sageattn_host: present, and never compute on CPU tensors."""
    import torch
    import sageattention_b200 as sab
    for name in ("QuantizedKV", "quantize_kv", "sageattn_prequantized", "sageattn_host"):
        assert hasattr(sab, name) and name in sab.__all__
    k = torch.zeros(1, 2, 64, 64, dtype=torch.float16)
    with pytest.raises(AssertionError):
        sab.quantize_kv(k, k)                     # CPU tensors are not silently computed
    with pytest.raises(ValueError):
        sab.sageattn_host(k, k, k, tensor_layout="XYZ")
    with pytest.raises(NotImplementedError):
        sab.sageattn_host(k, k, k, return_lse=True)
    with pytest.raises(AssertionError):
        sab.sageattn_host(k.float(), k.float(), k.float())
    from sageattention_b200.host import _chunks
    assert _chunks(2, 8, 2, 4) == [(0, 0, 4), (0, 4, 8), (1, 0, 4), (1, 4, 8)]
    assert _chunks(1, 6, 1, 4) == [(0, 0, 4), (0, 4, 6)]


def test_plain_c_program_links_and_validates(tmp_path):
    """tests/c/abi_smoke.c: a C (not C++) translation unit includes include/sageattn_b200.h, links libsageattn_b200.so with gcc
    alone and gets status codes + messages back — the boundary really is a C ABI with plain pointers and sizes."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from sageattention_b200 import _capi
    _capi.lib()
    libdir = os.path.dirname(_capi._LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    r = subprocess.run(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                        "-o", exe, "-L", libdir, "-lsageattn_b200", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "C_ABI_OK" in r.stdout, r.stdout + r.stderr

"""A/B harness for opt-in kernel experiments (compile-time variants of csrc/attn.cu next to the product library).

  python tools/ab_variants.py build [names...] # here (CPU box): nvcc the variants (default: all, 23 MB each — they travel with
                                               # the gpurun snapshot, so build only what the call will run) into lib/libsab_<name>.so
  python tools/ab_variants.py run [names...]   # on the GPU box: for the product library and each variant, the attention
                                               # parity tests + kernel-only timing (tools/perf_kernel.py), each in its own
                                               # subprocess under a timeout (a hanging variant costs 120 s, not the box)

The product library is never replaced: a variant is selected per process with SAB_LIB_PATH (sageattention_b200/_capi.py).
Variants whose P is not bit-identical to the reference kernel (polynomial exp2) run the parity tests with the tolerance
tests already state against the oracle (5e-3); the bit-exactness tests of S / P / m / d are expected to fail for them and
are excluded by -k."""
import os, re, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    # name: (defines, what it changes)
    "poly1": (["SAB_POLY_EXP_PAIRS=1"], "25 % of the exponentials as a degree-3 polynomial on the FMA pipe (ptx.cuh ex2_poly2)"),
    "poly2": (["SAB_POLY_EXP_PAIRS=2"], "50 % of the exponentials on the FMA pipe"),
    "premax8": (["SAB_PREMAX=8"], "row maxima of S(j+1) gathered inside the exp loop of tile j (8-column chunks)"),
    "defer": (["SAB_DEFER_PST"], "hand-off of P(j-1) (wait::st + fence + arrive) taken after S(j) was loaded, off the serial chain"),
    "late_alpha": (["SAB_LATE_ALPHA"], "alpha published after the first 8 exponentials of the tile instead of before the loop"),
    "all_chain": (["SAB_DEFER_PST", "SAB_PREMAX=8", "SAB_LATE_ALPHA"], "the three chain shorteners together (P stays bit-identical)"),
    "lazy3": (["SAB_LAZY_RESCALE=3"], "running max moves only when it grew by > 2^3: O rescale in ~3 % of the warp-tiles instead of ~80 %"),
    "all_chain_lazy3": (["SAB_DEFER_PST", "SAB_PREMAX=8", "SAB_LATE_ALPHA", "SAB_LAZY_RESCALE=3"], "chain shorteners + lazy rescale"),
    "lazy3_poly1": (["SAB_LAZY_RESCALE=3", "SAB_POLY_EXP_PAIRS=1"], "lazy rescale + 25 % polynomial exp2 (the hd64 kernel is issue-bound: fewer rescale instructions make room for the polynomial)"),
    "defer_premax8": (["SAB_DEFER_PST", "SAB_PREMAX=8"], "both chain shorteners"),
    "defer_premax8_poly1": (["SAB_DEFER_PST", "SAB_PREMAX=8", "SAB_POLY_EXP_PAIRS=1"], "chain shorteners + 25 % polynomial exp2"),
}
# attn_alt.cu (two softmax warpgroups on alternate key tiles) is selected at run time, on any library
ENVS = {
    "alt_wd": {"SAB_ATTN_KERNEL": "alt"},
    "alt": {"SAB_ATTN_KERNEL": "alt"},
    "alt_tau3": {"SAB_ATTN_KERNEL": "alt"},
    "alt_tau3_poly1": {"SAB_ATTN_KERNEL": "alt"},
}
VARIANTS.update({
    # watchdog build first: a wait that never completes reports itself (block 0) and is abandoned instead of hanging the GPU
    "alt_wd": (["SAB_WATCHDOG"], "attn_alt.cu with bounded mbarrier waits (attn_common.cuh mbar_wait_wd): run BEFORE the other alt variants"),
    "alt": ([], "attn_alt.cu with the exact max rule (in-line O rescale in most tiles)"),
    "alt_tau3": (["SAB_ALT_TAU=3"], "attn_alt.cu with the lazy max (tau = 3)"),
    "alt_tau3_poly1": (["SAB_ALT_TAU=3", "SAB_POLY_EXP_PAIRS=1"], "attn_alt.cu, lazy max, 25 % polynomial exp2"),
})
PARITY_K = "attention_vs_oracle or full_size_config1 or api_behaviour"


def build(names=()):
    from sageattention_b200 import build as b
    for name, (defs, what) in VARIANTS.items():
        if names and name not in names:
            continue
        t0 = time.time()
        if not defs:      # run-time variant of the product library: build it as a variant all the same (one path in run())
            defs = ["SAB_VARIANT_TAG=1"]
        lib = b.build(variant=name, defines=defs)
        print(f"[build] {name:16s} {time.time() - t0:5.1f} s  {lib}  ({what})", flush=True)


def _sub(cmd, env, tmo):
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=tmo)
        return r.returncode, r.stdout, r.stderr
    except subprocess.TimeoutExpired:
        return None, "", ""


def run(names, top_k=3):
    """Kernel timing for every variant first (cheap), then the parity subset only for the product and the top_k fastest
    variants (by the first shape: hd128 S=8192 non-causal per-thread)."""
    libdir = os.path.join(ROOT, "sageattention_b200", "lib")
    todo = [("product", None)] + [(n, os.path.join(libdir, f"libsab_{n}.so")) for n in (names or VARIANTS)]
    alt_ok = True
    score, envs = {}, {}
    for name, lib in todo:
        if name.startswith("alt") and name != "alt_wd" and not alt_ok:
            print(f"== {name}: skipped (alt_wd did not pass)", flush=True)
            continue
        if lib is not None and not os.path.exists(lib):
            print(f"== {name}: {lib} not built (python tools/ab_variants.py build)", flush=True)
            continue
        env = dict(os.environ)
        if lib:
            env["SAB_LIB_PATH"] = lib
        env.update(ENVS.get(name, {}))
        envs[name] = env
        rc, out, err = _sub([sys.executable, os.path.join(ROOT, "tools", "perf_kernel.py")], env, 150)
        if rc is None:
            print(f"== {name}: perf TIMEOUT (variant hangs?)", flush=True)
            if name == "alt_wd":
                alt_ok = False
            continue
        tail = (out.strip().splitlines() or [""])[-1]
        print(f"== {name}: perf rc={rc}: {tail}", flush=True)
        if name == "alt_wd" and (rc != 0 or "mbarrier timeout" in out + err):
            alt_ok = False
        if rc != 0:
            print("   " + "\n   ".join((out + err).strip().splitlines()[-12:]), flush=True)
            continue
        m = re.search(r"per_thread: (\d+)", tail)
        if m and name != "alt_wd":
            score[name] = int(m.group(1))
    best = sorted((n for n in score if n != "product"), key=lambda n: -score[n])[:top_k]
    for name in ["product"] + best:
        rc, out, err = _sub([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                             "-k", PARITY_K, "-p", "no:cacheprovider"], envs[name], 240)
        tail = (out.strip().splitlines() or [""])[-1] if rc is not None else "TIMEOUT"
        print(f"== {name}: parity rc={rc}: {tail}", flush=True)
        if rc not in (0, None):
            print("   " + "\n   ".join((out + err).strip().splitlines()[-25:]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) > 1 and sys.argv[1] == "run":
        run(sys.argv[2:])
    else:
        print(__doc__)

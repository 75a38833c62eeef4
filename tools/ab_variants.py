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
    # name: (defines, what it changes).  The round-2 experiments whose macros were deleted with the losing code paths
    # (polynomial exp2, premax, deferred hand-off, pair / split / speculative kernels) are recorded in profiles/r02_ab_*.log.
    "wd": (["SAB_WATCHDOG"], "bounded mbarrier waits (attn_common.cuh mbar_wait_wd): run FIRST after touching a barrier protocol"),
    "tau3": (["SAB_ALT_TAU=3"], "attn_alt.cu lazy running max with tau = 3 binades (more rescales, smaller P range)"),
    "tau5": (["SAB_ALT_TAU=5"], "attn_alt.cu lazy running max with tau = 5 binades"),
    "nostore": (["SAB_NO_TMA_STORE"], "direct per-row global stores in the epilogue instead of the staged TMA store"),
}
ENVS = {}   # per-variant environment (e.g. {"exact": {"SAB_ATTN_KERNEL": "exact"}} selects the exact-max kernels of any build)
ENVS["exact"] = {"SAB_ATTN_KERNEL": "exact"}
ENVS["q4"] = {"SAB_ATTN_KERNEL": "q4"}
VARIANTS["q4"] = ([], "attn_q4.cu: one CTA per SM, four softmax warpgroups, separate P buffers (head_dim 128) of the product build")
VARIANTS["exact"] = ([], "the exact-max kernels (attn.cu / attn_hd64.cu kLazy=false) of the product build")
PARITY_K = "attention_vs_oracle or full_size_config1 or api_behaviour"


def build(names=()):
    from sageattention_b200 import build as b
    for name, (defs, what) in VARIANTS.items():
        if names and name not in names:
            continue
        t0 = time.time()
        if not defs:      # run-time variant of the product library: nothing to build
            continue
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
    todo = [("product", None)] + [(n, os.path.join(libdir, f"libsab_{n}.so") if VARIANTS[n][0] else None) for n in (names or VARIANTS)]
    alt_ok = True
    score, envs = {}, {}
    for name, lib in todo:
        if name != "wd" and not alt_ok:
            print(f"== {name}: skipped (the watchdog build did not pass)", flush=True)
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
            if name == "wd":
                alt_ok = False
            continue
        tail = (out.strip().splitlines() or [""])[-1]
        print(f"== {name}: perf rc={rc}: {tail}", flush=True)
        if name == "wd" and (rc != 0 or "mbarrier timeout" in out + err):
            alt_ok = False
        if rc != 0:
            print("   " + "\n   ".join((out + err).strip().splitlines()[-12:]), flush=True)
            continue
        m = re.search(r"per_thread: (\d+)", tail)
        if m and name != "wd":
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

"""Build the sm_100a CUDA library in-tree: sageattention_b200/lib/libsageattn_b200.so.

One nvcc invocation, no torch headers (the boundary is a plain C ABI, include/sageattn_b200.h).
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os, subprocess, sys, hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libsageattn_b200.so")
SOURCES = ["attn.cu", "attn_pair.cu", "attn_hd64.cu", "attn_split.cu", "quant.cu", "capi.cu"]
HEADERS = ["ptx.cuh", "common.cuh", "attn_common.cuh", os.path.join("..", "..", "include", "sageattn_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = LIB + ".sha256"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsageattn_b200.so")
    with open(os.path.join(os.path.dirname(LIB), "ptxas.log"), "w") as fh:
        fh.write(res.stdout + res.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

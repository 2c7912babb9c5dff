"""Build libfa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfa_b200.so")
SOURCES = ["fa_api.cu", "fa_bwd_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v",
]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def needs_build():
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_source_mtime()


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the library.  `out`/`extra_flags` build an experiment variant next to the default .so
    (selected at load time with FA_B200_LIB=<path>)."""
    out = out or LIB
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if os.environ.get("FA_B200_SPIN_GUARD") == "1":
        extra_flags = tuple(extra_flags) + ("-DFA_SPIN_GUARD=1",)
    cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(out + ".build.log" if out != LIB else os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose:
        print(log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)

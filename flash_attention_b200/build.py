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


STAMP = LIB + ".srchash"   # content hash of the sources the .so was built from (travels with it to the GPU box)


def source_hash():
    import hashlib

    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    """Stale = built from different source CONTENT (mtimes do not survive the copy to a GPU box)."""
    if not os.path.exists(LIB):
        return True
    try:
        return open(STAMP).read().strip() != source_hash()
    except OSError:
        return True


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
    if out == LIB and not extra_flags:
        with open(STAMP, "w") as f:
            f.write(source_hash() + "\n")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)

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
    # one SONAME for the default build and every experiment variant: the operator module (ext/) links against
    # "libfa_b200.so", and a variant pre-loaded by _lib.load() (FA_B200_LIB=...) then satisfies that dependency
    "-Xlinker", "-soname=libfa_b200.so",
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


# ------------------------------------------------------------------------------------------------------------------
# The compiled operator module `flash_attn_2_cuda` (pybind11 + TORCH_LIBRARY(flash_attn_3)): ATen-facing host code in
# csrc/ext/*.cpp, linked against libfa_b200.so (rpath $ORIGIN/..).  Plain g++ with torch's include / library paths -- the
# same flags torch.utils.cpp_extension would pass -- so that the result lands IN-TREE (it travels to the GPU box with
# the snapshot; a JIT cache under ~/.cache would not).
EXT_DIR = os.path.join(HERE, "ext")
EXT_SOURCES = ["flash_attn_2_cuda.cpp", "flash_attn_3_ops.cpp"]
EXT_STAMP = os.path.join(EXT_DIR, "flash_attn_2_cuda.srchash")


def ext_path():
    import sysconfig

    return os.path.join(EXT_DIR, "flash_attn_2_cuda" + sysconfig.get_config_var("EXT_SUFFIX"))


def ext_source_hash():
    import hashlib

    import torch

    h = hashlib.sha256(torch.__version__.encode())
    d = os.path.join(CSRC, "ext")
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode() + fh.read())
    with open(os.path.join(HERE, "..", "include", "fa_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def ext_needs_build():
    if not os.path.exists(ext_path()):
        return True
    try:
        return open(EXT_STAMP).read().strip() != ext_source_hash()
    except OSError:
        return True


def build_ext(force=False, verbose=False):
    """Compile the operator module (needs libfa_b200.so to exist: it is linked, not dlopen'ed)."""
    if not force and not ext_needs_build():
        return ext_path()
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce

    build(force=False)
    os.makedirs(EXT_DIR, exist_ok=True)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ce.include_paths() + [os.path.join(cuda_home, "include"), sysconfig.get_paths()["include"]]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    out = ext_path()
    objs = []
    cxx = os.environ.get("CXX", "g++")
    common = [cxx, "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=flash_attn_2_cuda",
              "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
              "-DPYBIND11_COMPILER_TYPE=\"_gcc\"", "-DPYBIND11_STDLIB=\"_libstdcpp\"", "-DPYBIND11_BUILD_ABI=\"_cxxabi1018\"",
              "-Wno-deprecated-declarations"] + [f"-I{i}" for i in inc]
    import concurrent.futures as cf

    def one(src):
        obj = os.path.join(EXT_DIR, src.replace(".cpp", ".o"))
        r = subprocess.run(common + ["-c", os.path.join(CSRC, "ext", src), "-o", obj], capture_output=True, text=True)
        return obj, r

    with cf.ThreadPoolExecutor(len(EXT_SOURCES)) as ex:
        for obj, r in ex.map(one, EXT_SOURCES):
            if verbose or r.returncode != 0:
                print(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("g++ failed:\n" + (r.stdout + r.stderr)[-6000:])
            objs.append(obj)
    link = [cxx, "-shared", "-o", out] + objs + [f"-L{torch_lib}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", "-lc10_cuda",
            "-ltorch_cuda", f"-L{os.path.join(cuda_home, 'lib64')}", "-lcudart", f"-L{HERE}", "-lfa_b200",
            "-Wl,-rpath,$ORIGIN/..", f"-Wl,-rpath,{torch_lib}", f"-Wl,-rpath,{os.path.join(cuda_home, 'lib64')}"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + (r.stdout + r.stderr)[-4000:])
    for o in objs:
        os.remove(o)
    with open(EXT_STAMP, "w") as f:
        f.write(ext_source_hash() + "\n")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
    print("built", build_ext(force="--force" in sys.argv, verbose=True))

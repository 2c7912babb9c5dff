"""`flash_attn_2_cuda` — loader of the COMPILED operator module (csrc/ext/flash_attn_2_cuda.cpp + flash_attn_3_ops.cpp).

The module has the exact surface of the reference's extension (csrc/flash_attn/flash_api.cpp:1535-1542): `fwd`,
`varlen_fwd`, `bwd`, `varlen_bwd`, `fwd_kvcache` with positional arguments, each returning a list of tensors, so the
reference's `flash_attn/flash_attn_interface.py` runs unchanged on top of it:

    import flash_attention_b200; flash_attention_b200.install()   # sys.modules['flash_attn_2_cuda'] = the compiled module
    from flash_attn.flash_attn_interface import flash_attn_func   # the reference's own file

All host logic (argument checks, allocation, dropout state, dispatch) is C++ in that module; the compute goes through the
C ABI of libfa_b200.so (include/fa_b200.h).  This file only locates / builds / imports the shared object and re-exports
its functions (plus the extras `fwd_fp8`, `fwd_ex`, `fwd_combine`, `plan_kvcache`).  There is no Python fallback: if the
module cannot be built or loaded, importing this file raises.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import torch  # noqa: F401  (libtorch must be loaded before the extension)

from . import _lib
from . import build as _build

_MOD = None


def load():
    """Import the compiled module (building it first when it is missing or its sources changed)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    _lib.load()  # libfa_b200.so (or the FA_B200_LIB variant) first, RTLD_GLOBAL: the extension links against its SONAME
    path = _build.ext_path()
    if not os.path.exists(path) or os.environ.get("FA_B200_REBUILD") == "1":
        _build.build_ext(force=True)
    elif _build.ext_needs_build():
        try:
            _build.build_ext()
        except Exception as e:  # noqa: BLE001  (a GPU box ships the prebuilt module; say so instead of hiding it)
            import warnings

            warnings.warn(f"flash_attn_2_cuda extension is older than its sources and could not be rebuilt ({e}); loading it as is")
    # The shared object's init symbol is PyInit_flash_attn_2_cuda (put flash_attention_b200/ext on sys.path and a plain
    # `import flash_attn_2_cuda` works).  Here it is imported under a qualified spec name: pybind11 >= 3 caches extension
    # modules per interpreter BY SPEC NAME, so with the reference's own flash_attn_2_cuda wheel already imported in the
    # same process (the parity tests run both side by side) a second spec called "flash_attn_2_cuda" would silently
    # return the reference's module object instead of initialising this one.
    spec = importlib.util.spec_from_file_location("flash_attention_b200.ext.flash_attn_2_cuda", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _MOD = mod
    return mod


_m = load()
fwd = _m.fwd
varlen_fwd = _m.varlen_fwd
bwd = _m.bwd
varlen_bwd = _m.varlen_bwd
fwd_kvcache = _m.fwd_kvcache
fwd_fp8 = _m.fwd_fp8
fwd_ex = _m.fwd_ex
fwd_combine = _m.fwd_combine
plan_kvcache = _m.plan_kvcache
launch_count = _m.launch_count
compiled_module = _m

"""flash_attention_b200 — B200-native (sm_100a) fused attention behind the flash_attn operator API.

    from flash_attention_b200 import flash_attn_func, flash_attn_varlen_func, flash_attn_with_kvcache

or, to run the reference's own `flash_attn/flash_attn_interface.py` unchanged on these kernels:

    import flash_attention_b200; flash_attention_b200.install()
"""
import sys

__version__ = "0.1.0"


def install(name: str = "flash_attn_2_cuda"):
    """Register this package's COMPILED operator module as `flash_attn_2_cuda` (the extension module the
    reference's flash_attn_interface.py imports, flash_attn/flash_attn_interface.py:23)."""
    from . import flash_attn_2_cuda as loader

    sys.modules[name] = loader.compiled_module
    return loader.compiled_module


def install_fa3():
    """Make `import flash_attn_3._C` (hopper/flash_attn_interface.py:24) resolve to this package's compiled module, whose
    import registers TORCH_LIBRARY(flash_attn_3) {fwd, bwd, fwd_combine, get_scheduler_metadata} (csrc/ext/flash_attn_3_ops.cpp),
    so that the reference's FA3 interface file runs unchanged on these kernels via `torch.ops.flash_attn_3`."""
    import types

    from . import flash_attn_2_cuda as loader

    pkg = sys.modules.get("flash_attn_3")
    if pkg is None:
        pkg = types.ModuleType("flash_attn_3")
        pkg.__path__ = []
        sys.modules["flash_attn_3"] = pkg
    pkg._C = loader.compiled_module
    sys.modules["flash_attn_3._C"] = loader.compiled_module
    return loader.compiled_module


def __getattr__(attr):
    # lazy: importing the package must not require torch/CUDA (build() runs on a CPU box)
    if attr in ("flash_attn_func", "flash_attn_varlen_func", "flash_attn_with_kvcache", "flash_attn_qkvpacked_func",
                "flash_attn_kvpacked_func", "flash_attn_varlen_qkvpacked_func", "flash_attn_varlen_kvpacked_func",
                "flash_attn_combine"):
        from . import flash_attn_interface as m

        return getattr(m, attr)
    if attr == "flash_attn_3":  # FA3-style surface: flash_attention_b200.flash_attn_3.flash_attn_func(...)
        from . import flash_attn_3_interface as m3

        return m3
    raise AttributeError(attr)

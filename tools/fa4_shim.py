"""Compatibility shim that lets the UNMODIFIED reference FA4 (flash_attn/cute, installed into baseline/_ref) trace with the
image's nvidia-cutlass-dsl 4.5.0 (FA4 declares >= 4.6.2, flash_attn/cute/pyproject.toml:27-35; no network to upgrade).
Only the DSL's Python bindings are wrapped -- nothing under baseline/_ref is edited.  Used by tools/baseline_probe.py to get
a same-box Blackwell comparator; every patch says which call site needed it.

    import tools.fa4_shim as shim; shim.apply()            # before importing flash_attn.cute
    shim.apply(cpu_probe=True)                             # compile-only on a CPU box (stubs the device query)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APPLIED = []


def _patch_nvvm_fmax():
    # flash_attn/cute/utils.py:350-362 calls nvvm.fmax(a, b, c=...) (DSL >= 4.6 infers the result type); 4.5.0's
    # binding is fmax(res, a, b, *, c=...): insert res = a.type when the first argument is a Value, not a Type.
    from cutlass._mlir import ir
    from cutlass._mlir.dialects import nvvm

    for name in ("fmax", "fmin"):
        orig = getattr(nvvm, name, None)
        if orig is None or getattr(orig, "_fa4_shim", False):
            continue

        def wrapped(*args, _orig=orig, **kw):
            if args and isinstance(args[0], ir.Value):
                return _orig(args[0].type, *args, **kw)
            return _orig(*args, **kw)

        wrapped._fa4_shim = True
        setattr(nvvm, name, wrapped)
        APPLIED.append(f"nvvm.{name}: result type inserted (flash_attn/cute/utils.py:350-362)")


def _stub_hardware_info():
    # CPU-only tracing: tile_scheduler.py:409 asks cutlass.utils.HardwareInfo for the SM count through the driver
    import cutlass.utils as cu

    class _HW:
        def __init__(self, *a, **k):
            pass

        def get_max_active_clusters(self, cluster_size, *a, **k):
            return 148 // max(1, int(cluster_size))

        def get_device_multiprocessor_count(self, *a, **k):
            return 148

        def get_l2_cache_size_in_bytes(self, *a, **k):
            return 126 * 1024 * 1024

    cu.HardwareInfo = _HW
    try:
        import cutlass.utils.hardware_info as hi

        hi.HardwareInfo = _HW
    except Exception:  # noqa: BLE001
        pass
    APPLIED.append("cutlass.utils.HardwareInfo stubbed (CPU probe only)")


def apply(cpu_probe=False):
    _patch_nvvm_fmax()
    if cpu_probe:
        _stub_hardware_info()
    import flash_attn

    p = os.path.join(ROOT, "baseline", "_ref", "flash_attn")
    if p not in flash_attn.__path__:
        flash_attn.__path__.append(p)
    return APPLIED

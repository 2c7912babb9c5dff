"""ctypes binding of libfa_b200.so (the C ABI declared in include/fa_b200.h).

The product path has NO fallback: if the library cannot be built/loaded the import of the
operator module raises.  (The CPU oracle under oracle/ is test infrastructure and is never
imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

FP16, BF16, FP8_E4M3 = 0, 1, 2
ENTRY_INFER, ENTRY_DENSE, ENTRY_VARLEN, ENTRY_KVCACHE = 0, 1, 2, 3

_i64, _i32, _f32, _vp, _sz = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_size_t


class FwdArgs(C.Structure):
    """Mirror of `fa_b200_fwd_args` (include/fa_b200.h) — field order is the ABI."""

    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("out", _vp), ("softmax_lse", _vp),
        ("q_batch_stride", _i64), ("q_row_stride", _i64), ("q_head_stride", _i64),
        ("k_batch_stride", _i64), ("k_row_stride", _i64), ("k_head_stride", _i64),
        ("v_batch_stride", _i64), ("v_row_stride", _i64), ("v_head_stride", _i64),
        ("o_batch_stride", _i64), ("o_row_stride", _i64), ("o_head_stride", _i64),
        ("b", _i32), ("seqlen_q", _i32), ("seqlen_k", _i32), ("h", _i32), ("h_k", _i32), ("d", _i32),
        ("total_q", _i32), ("total_k", _i32),
        ("dtype", _i32), ("b_kvcache", _i32),
        ("cu_seqlens_q", _vp), ("cu_seqlens_k", _vp), ("seqused_k", _vp), ("seqused_q", _vp),
        ("cache_batch_idx", _vp), ("leftpad_k", _vp), ("block_table", _vp),
        ("block_table_batch_stride", _i64),
        ("page_block_size", _i32), ("num_pages", _i32),
        ("alibi_slopes", _vp), ("alibi_slopes_batch_stride", _i64),
        ("softmax_scale", _f32), ("softcap", _f32),
        ("is_causal", _i32), ("window_size_left", _i32), ("window_size_right", _i32),
        ("q_descale", _vp), ("k_descale", _vp), ("v_descale", _vp),
        ("q_descale_batch_stride", _i64), ("k_descale_batch_stride", _i64), ("v_descale_batch_stride", _i64),
        ("num_splits", _i32), ("pack_gqa", _i32),
        ("workspace", _vp), ("workspace_bytes", _sz),
        ("debug_s", _vp),
        ("p_dropout", C.c_float), ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
        ("s_dmask", _vp),
        ("entry", _i32), ("sm_margin", _i32),
    ]


class BwdArgs(C.Structure):
    """Mirror of `fa_b200_bwd_args`."""

    _fields_ = [
        ("dout", _vp), ("q", _vp), ("k", _vp), ("v", _vp), ("out", _vp), ("softmax_lse", _vp),
        ("dq", _vp), ("dk", _vp), ("dv", _vp), ("softmax_d", _vp),
        ("q_batch_stride", _i64), ("q_row_stride", _i64), ("q_head_stride", _i64),
        ("k_batch_stride", _i64), ("k_row_stride", _i64), ("k_head_stride", _i64),
        ("v_batch_stride", _i64), ("v_row_stride", _i64), ("v_head_stride", _i64),
        ("o_batch_stride", _i64), ("o_row_stride", _i64), ("o_head_stride", _i64),
        ("do_batch_stride", _i64), ("do_row_stride", _i64), ("do_head_stride", _i64),
        ("dq_batch_stride", _i64), ("dq_row_stride", _i64), ("dq_head_stride", _i64),
        ("dk_batch_stride", _i64), ("dk_row_stride", _i64), ("dk_head_stride", _i64),
        ("dv_batch_stride", _i64), ("dv_row_stride", _i64), ("dv_head_stride", _i64),
        ("b", _i32), ("seqlen_q", _i32), ("seqlen_k", _i32), ("h", _i32), ("h_k", _i32), ("d", _i32),
        ("total_q", _i32), ("total_k", _i32),
        ("dtype", _i32),
        ("cu_seqlens_q", _vp), ("cu_seqlens_k", _vp),
        ("alibi_slopes", _vp), ("alibi_slopes_batch_stride", _i64),
        ("softmax_scale", _f32), ("softcap", _f32),
        ("is_causal", _i32), ("window_size_left", _i32), ("window_size_right", _i32),
        ("deterministic", _i32),
        ("workspace", _vp), ("workspace_bytes", _sz),
        ("p_dropout", C.c_float), ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
        ("rng_state", _vp), ("sm_margin", _i32),
    ]


class AppendArgs(C.Structure):
    """Mirror of `fa_b200_append_args`."""

    _fields_ = [
        ("kcache", _vp), ("vcache", _vp), ("knew", _vp), ("vnew", _vp), ("q", _vp), ("q_out", _vp),
        ("kc_batch_stride", _i64), ("kc_row_stride", _i64), ("kc_head_stride", _i64),
        ("vc_batch_stride", _i64), ("vc_row_stride", _i64), ("vc_head_stride", _i64),
        ("kn_batch_stride", _i64), ("kn_row_stride", _i64), ("kn_head_stride", _i64),
        ("vn_batch_stride", _i64), ("vn_row_stride", _i64), ("vn_head_stride", _i64),
        ("q_batch_stride", _i64), ("q_row_stride", _i64), ("q_head_stride", _i64),
        ("qo_batch_stride", _i64), ("qo_row_stride", _i64), ("qo_head_stride", _i64),
        ("cache_seqlens", _vp), ("cache_batch_idx", _vp), ("leftpad_k", _vp),
        ("block_table", _vp), ("block_table_batch_stride", _i64), ("page_block_size", _i32),
        ("rotary_cos", _vp), ("rotary_sin", _vp),
        ("rotary_dim", _i32), ("is_rotary_interleaved", _i32), ("q_pos_per_row", _i32),
        ("b", _i32), ("seqlen_new", _i32), ("seqlen_q", _i32), ("h", _i32), ("h_k", _i32), ("d", _i32), ("dtype", _i32),
    ]


# every symbol include/fa_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "fa_b200_mha_fwd": (C.c_int, [C.POINTER(FwdArgs), _vp]),
    "fa_b200_mha_varlen_fwd": (C.c_int, [C.POINTER(FwdArgs), _vp]),
    "fa_b200_mha_fwd_kvcache": (C.c_int, [C.POINTER(FwdArgs), _vp]),
    "fa_b200_fwd_workspace_bytes": (_sz, [C.POINTER(FwdArgs)]),
    "fa_b200_fwd_num_splits": (C.c_int, [C.POINTER(FwdArgs)]),
    "fa_b200_fwd_kernel_kind": (C.c_int, [C.POINTER(FwdArgs)]),
    "fa_b200_fwd_plan": (C.c_int, [C.POINTER(FwdArgs), C.POINTER(C.c_int32)]),
    "fa_b200_mha_bwd": (C.c_int, [C.POINTER(BwdArgs), _vp]),
    "fa_b200_mha_varlen_bwd": (C.c_int, [C.POINTER(BwdArgs), _vp]),
    "fa_b200_bwd_workspace_bytes": (_sz, [C.POINTER(BwdArgs)]),
    "fa_b200_fwd_combine": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp]),
    "fa_b200_kvcache_append": (C.c_int, [C.POINTER(AppendArgs), _vp]),
    "fa_b200_sizeof_append_args": (_sz, []),
    "fa_b200_memcpy2d_async": (C.c_int, [_vp, _sz, _vp, _sz, _sz, _sz, _i32, _vp]),
    "fa_b200_write_rng_state": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "fa_b200_num_splits_heuristic": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fa_b200_last_error": (C.c_char_p, []),
    "fa_b200_sizeof_fwd_args": (_sz, []),
    "fa_b200_sizeof_bwd_args": (_sz, []),
    "fa_b200_version": (C.c_char_p, []),
    "fa_b200_launch_count": (C.c_uint64, []),
}


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the .so is missing or stale). Raises on any failure — no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("FA_B200_LIB") or _build.LIB  # FA_B200_LIB: experiment variant built by build.build(out=...)
    if path == _build.LIB:
        if not os.path.exists(path) or os.environ.get("FA_B200_REBUILD") == "1":
            _build.build(force=True)
        elif _build.needs_build():   # a source is newer than the .so: rebuild when a compiler is here, else say so
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001  (GPU boxes ship the prebuilt .so; mtimes of a fresh copy may be unordered)
                import warnings

                warnings.warn(f"libfa_b200.so is older than its sources and could not be rebuilt ({e}); loading it as is")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)   # global: the compiled operator module (ext/) resolves its libfa_b200.so dependency to THIS object
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if (lib.fa_b200_sizeof_fwd_args() != C.sizeof(FwdArgs) or lib.fa_b200_sizeof_bwd_args() != C.sizeof(BwdArgs)
            or lib.fa_b200_sizeof_append_args() != C.sizeof(AppendArgs)):
        raise RuntimeError(
            "fa_b200 ABI mismatch: ctypes struct size %d/%d vs library %d/%d"
            % (C.sizeof(FwdArgs), C.sizeof(BwdArgs), lib.fa_b200_sizeof_fwd_args(), lib.fa_b200_sizeof_bwd_args())
        )
    _LIB = lib
    return lib


def last_error() -> str:
    return load().fa_b200_last_error().decode()


def check(rc: int):
    if rc != 0:
        raise RuntimeError(last_error())

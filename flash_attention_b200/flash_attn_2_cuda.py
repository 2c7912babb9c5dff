"""Operator module with the exact surface of the reference's `flash_attn_2_cuda` extension
(csrc/flash_attn/flash_api.cpp:1535-1542): `fwd`, `varlen_fwd`, `bwd`, `varlen_bwd`, `fwd_kvcache`,
positional arguments, each returning a list of tensors.  The reference's
`flash_attn/flash_attn_interface.py` runs unchanged on top of it:

    import flash_attention_b200; flash_attention_b200.install()   # registers sys.modules['flash_attn_2_cuda']
    from flash_attn.flash_attn_interface import flash_attn_func   # the reference's own file

All compute goes through the C ABI of libfa_b200.so (include/fa_b200.h); this file only checks
arguments the way flash_api.cpp does, allocates outputs/scratch with torch and passes raw device
pointers + element strides + the current CUDA stream.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional

import torch

from . import _lib

_DT = {torch.float16: _lib.FP16, torch.bfloat16: _lib.BF16}
_PACK = {None: 0, True: 1, False: 2}  # FA3 `pack_gqa` -> fa_b200_fwd_args.pack_gqa


def _check(cond: bool, msg: str):
    if not cond:
        raise RuntimeError(msg)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _common_checks(q, k, v):
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "Input tensor must be on CUDA device")
    _check(q.dtype in _DT, "FlashAttention only support fp16 and bf16 data type")
    _check(k.dtype == q.dtype, "query and key must have the same dtype")
    _check(v.dtype == q.dtype, "query and value must have the same dtype")
    _check(q.stride(-1) == 1, "Input tensor must have contiguous last dimension")
    _check(k.stride(-1) == 1, "Input tensor must have contiguous last dimension")
    _check(v.stride(-1) == 1, "Input tensor must have contiguous last dimension")


def _alibi(args, alibi_slopes, b, h):
    if alibi_slopes is None:
        return
    _check(alibi_slopes.dtype == torch.float32, "ALiBi slopes must have dtype fp32")
    _check(alibi_slopes.is_cuda, "Input tensor must be on CUDA device")
    _check(alibi_slopes.stride(-1) == 1, "ALiBi slopes tensor must have contiguous last dimension")
    _check(tuple(alibi_slopes.shape) in ((h,), (b, h)), "ALiBi slopes must have shape (nheads,) or (batch, nheads)")
    args.alibi_slopes = _ptr(alibi_slopes)
    args.alibi_slopes_batch_stride = alibi_slopes.stride(0) if alibi_slopes.dim() == 2 else 0


def _workspace(lib, args, device):
    nbytes = lib.fa_b200_fwd_workspace_bytes(C.byref(args))
    if nbytes == 0:
        return None
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    args.workspace = _ptr(ws)
    args.workspace_bytes = nbytes
    return ws


_ZERO_RNG_STATE = {}


def _round128(x: int) -> int:
    return (x + 127) // 128 * 128


def _dropout_state(a, p_dropout, gen_, device, b, h):
    """Philox (seed, offset) of this call from the CUDA generator, advanced like the reference does
    (flash_api.cpp:505-517: counter_offset = b * h * 32).  Returns the `rng_state` tensor (2,) int64."""
    _check(0.0 <= p_dropout < 1.0, "p_dropout must be in [0, 1)")
    if p_dropout == 0.0:  # one shared (never written) zero state per device instead of a fill kernel per call
        z = _ZERO_RNG_STATE.get(device)
        if z is None:
            z = _ZERO_RNG_STATE[device] = torch.zeros(2, dtype=torch.int64, device=device)
        return z
    gen = gen_ if gen_ is not None else torch.cuda.default_generators[device.index if device.index is not None
                                                                       else torch.cuda.current_device()]
    seed, offset = int(gen.initial_seed()), int(gen.get_offset())
    gen.set_offset(offset + b * h * 32)
    a.p_dropout, a.philox_seed, a.philox_offset = p_dropout, seed & (2**64 - 1), offset
    wrap = lambda x: x - 2**64 if x >= 2**63 else x  # noqa: E731  (int64 storage of the uint64 seed)
    return torch.tensor([wrap(seed & (2**64 - 1)), offset], dtype=torch.int64).to(device, non_blocking=True)


# --------------------------------------------------------------------------------------- fwd
def fwd(q, k, v, out_, alibi_slopes_, p_dropout, softmax_scale, is_causal, window_size_left, window_size_right,
        softcap, return_softmax, gen_=None) -> List[torch.Tensor]:
    """mha_fwd (flash_api.cpp:368-536).  q (b,sq,h,d), k/v (b,sk,h_k,d) -> [out, softmax_lse(b,h,sq), p, rng_state]."""
    lib = _lib.load()
    _common_checks(q, k, v)
    if return_softmax:
        _check(p_dropout > 0.0, "return_softmax is only supported when p_dropout > 0.0")
    _check(q.dim() == 4 and k.dim() == 4 and v.dim() == 4, "q, k, v must be 4-D (batch, seqlen, nheads, headdim)")
    b, sq, h, d = q.shape
    sk, hk = k.shape[1], k.shape[2]
    _check(b > 0, "batch size must be positive")
    _check(d <= 256, "FlashAttention forward only supports head dimension at most 256")
    _check(d % 8 == 0, "query, key, value, and out_ must have a head_size that is a multiple of 8")
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    _check(tuple(k.shape) == (b, sk, hk, d), "k has the wrong shape")
    _check(tuple(v.shape) == (b, sk, hk, d), "v has the wrong shape")
    if softcap > 0.0:
        _check(p_dropout == 0.0, "Softcapping does not support dropout for now")
    if out_ is not None:
        out = out_
        _check(out.dtype == q.dtype, "Output must have the same dtype as inputs")
        _check(out.is_cuda and out.stride(-1) == 1, "Output tensor must have contiguous last dimension")
        _check(tuple(out.shape) == (b, sq, h, d), "out has the wrong shape")
    else:
        out = torch.empty_like(q)
    lse = torch.empty((b, h, sq), dtype=torch.float32, device=q.device)
    p = torch.empty(0, dtype=q.dtype, device=q.device)
    a = _lib.FwdArgs()
    rng_state = _dropout_state(a, p_dropout, gen_, q.device, b, h)
    if sq == 0:
        return [out, lse, p, rng_state]
    if sk == 0:  # flash_api.cpp:521-528
        out.zero_()
        lse.fill_(float("inf"))
        return [out, lse, p, rng_state]
    if return_softmax:  # (b, h, seqlen_q_rounded, seqlen_k_rounded), flash_api.cpp:466-470
        p = torch.empty((b, h, _round128(sq), _round128(sk)), dtype=q.dtype, device=q.device)
        a.s_dmask = _ptr(p)
    a.q, a.k, a.v, a.out, a.softmax_lse = _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse)
    a.q_batch_stride, a.q_row_stride, a.q_head_stride = q.stride(0), q.stride(1), q.stride(2)
    a.k_batch_stride, a.k_row_stride, a.k_head_stride = k.stride(0), k.stride(1), k.stride(2)
    a.v_batch_stride, a.v_row_stride, a.v_head_stride = v.stride(0), v.stride(1), v.stride(2)
    a.o_batch_stride, a.o_row_stride, a.o_head_stride = out.stride(0), out.stride(1), out.stride(2)
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, sq, sk, h, hk, d
    a.dtype = _DT[q.dtype]
    a.softmax_scale, a.softcap = softmax_scale, softcap
    a.is_causal, a.window_size_left, a.window_size_right = int(is_causal), window_size_left, window_size_right
    a.num_splits = 0
    _alibi(a, alibi_slopes_, b, h)
    with torch.cuda.device(q.device):
        ws = _workspace(lib, a, q.device)  # noqa: F841 (kept alive until the launch is enqueued)
        _lib.check(lib.fa_b200_mha_fwd(C.byref(a), _stream(q)))
    return [out, lse, p, rng_state]


# --------------------------------------------------------------------------------------- fwd_fp8
def fwd_fp8(q, k, v, out_, q_descale, k_descale, v_descale, softmax_scale, is_causal, window_size_left,
            window_size_right, softcap=0.0, alibi_slopes_=None, num_splits=0, pack_gqa=None) -> List[torch.Tensor]:
    """FP8 e4m3 forward (no FA2 equivalent: flash_api.cpp:396 rejects fp8; spec = FA3 `mha_fwd` descale
    arguments, hopper/flash_api.cpp, and flash_attn/cute/interface.py:637-643).  q/k/v torch.float8_e4m3fn,
    descales (b, h_k) fp32 or None (= 1.0); output bf16 -> [out, softmax_lse(b,h,sq)]."""
    lib = _lib.load()
    f8 = torch.float8_e4m3fn
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "Input tensor must be on CUDA device")
    _check(q.dtype == f8 and k.dtype == f8 and v.dtype == f8, "fwd_fp8 expects float8_e4m3fn q, k, v")
    _check(q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1, "Input tensor must have contiguous last dimension")
    b, sq, h, d = q.shape
    sk, hk = k.shape[1], k.shape[2]
    _check(d % 16 == 0 and d <= 128, "fp8 forward supports head dimensions that are multiples of 16 up to 128")
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    out = out_ if out_ is not None else torch.empty((b, sq, h, d), dtype=torch.bfloat16, device=q.device)
    _check(out.dtype == torch.bfloat16 and out.stride(-1) == 1 and tuple(out.shape) == (b, sq, h, d), "out must be bf16 (b,sq,h,d)")
    lse = torch.empty((b, h, sq), dtype=torch.float32, device=q.device)
    if sq == 0:
        return [out, lse]
    if sk == 0:
        out.zero_()
        lse.fill_(float("inf"))
        return [out, lse]
    a = _lib.FwdArgs()
    a.q, a.k, a.v, a.out, a.softmax_lse = _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse)
    a.q_batch_stride, a.q_row_stride, a.q_head_stride = q.stride(0), q.stride(1), q.stride(2)
    a.k_batch_stride, a.k_row_stride, a.k_head_stride = k.stride(0), k.stride(1), k.stride(2)
    a.v_batch_stride, a.v_row_stride, a.v_head_stride = v.stride(0), v.stride(1), v.stride(2)
    a.o_batch_stride, a.o_row_stride, a.o_head_stride = out.stride(0), out.stride(1), out.stride(2)
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, sq, sk, h, hk, d
    a.dtype = _lib.FP8_E4M3
    keep = []
    for name, t in (("q_descale", q_descale), ("k_descale", k_descale), ("v_descale", v_descale)):
        if t is not None:
            _check(t.dtype == torch.float32 and t.is_cuda and tuple(t.shape) == (b, hk), f"{name} must be fp32 (batch, nheads_k)")
            t = t.contiguous()
            keep.append(t)
            setattr(a, name, _ptr(t))
            setattr(a, name + "_batch_stride", t.stride(0))
    a.softmax_scale, a.softcap = softmax_scale, softcap
    a.is_causal, a.window_size_left, a.window_size_right = int(is_causal), window_size_left, window_size_right
    a.num_splits = num_splits
    a.pack_gqa = _PACK[pack_gqa]
    _alibi(a, alibi_slopes_, b, h)
    with torch.cuda.device(q.device):
        ws = _workspace(lib, a, q.device)  # noqa: F841
        _lib.check(lib.fa_b200_mha_fwd(C.byref(a), _stream(q)))
    return [out, lse]


def fwd_ex(q, k, v, softmax_scale, is_causal, window_size_left, window_size_right, softcap=0.0, num_splits=0,
           pack_gqa=None, alibi_slopes_=None) -> List[torch.Tensor]:
    """Dense forward with the FA3 scheduling knobs (`num_splits`, `pack_gqa`; hopper/flash_api.cpp mha_fwd) -> [out, lse]."""
    lib = _lib.load()
    _common_checks(q, k, v)
    b, sq, h, d = q.shape
    sk, hk = k.shape[1], k.shape[2]
    _check(d % 8 == 0 and h % hk == 0, "bad head_size / number of heads")
    out = torch.empty_like(q)
    lse = torch.empty((b, h, sq), dtype=torch.float32, device=q.device)
    a = _lib.FwdArgs()
    a.q, a.k, a.v, a.out, a.softmax_lse = _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse)
    a.q_batch_stride, a.q_row_stride, a.q_head_stride = q.stride(0), q.stride(1), q.stride(2)
    a.k_batch_stride, a.k_row_stride, a.k_head_stride = k.stride(0), k.stride(1), k.stride(2)
    a.v_batch_stride, a.v_row_stride, a.v_head_stride = v.stride(0), v.stride(1), v.stride(2)
    a.o_batch_stride, a.o_row_stride, a.o_head_stride = out.stride(0), out.stride(1), out.stride(2)
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, sq, sk, h, hk, d
    a.dtype = _DT[q.dtype]
    a.softmax_scale, a.softcap = softmax_scale, softcap
    a.is_causal, a.window_size_left, a.window_size_right = int(is_causal), window_size_left, window_size_right
    a.num_splits, a.pack_gqa = num_splits, _PACK[pack_gqa]
    _alibi(a, alibi_slopes_, b, h)
    with torch.cuda.device(q.device):
        ws = _workspace(lib, a, q.device)  # noqa: F841
        _lib.check(lib.fa_b200_mha_fwd(C.byref(a), _stream(q)))
    return [out, lse]


def plan_kvcache(batch, seqlen_q, seqlen_k, h, hk, d, page_size, causal, wl, wr, num_splits=0, pack_gqa=None):
    """Host-only: (num_splits, gqa_pack_factor, num_m_blocks) the library will use for this decode shape."""
    lib = _lib.load()
    a = _lib.FwdArgs()
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = batch, seqlen_q, seqlen_k, h, hk, d
    a.dtype = _lib.BF16
    a.seqused_k = C.c_void_p(16)  # non-null marks the kvcache mode for the planner; never dereferenced on the host
    a.is_causal, a.window_size_left, a.window_size_right = int(causal), wl, wr
    a.num_splits, a.pack_gqa = num_splits, _PACK[pack_gqa]
    if page_size:
        a.block_table, a.page_block_size, a.num_pages = C.c_void_p(16), page_size, 1
    out = (C.c_int32 * 3)()
    _lib.check(lib.fa_b200_fwd_plan(C.byref(a), out))
    return int(out[0]), int(out[1]), int(out[2])


# --------------------------------------------------------------------------------------- varlen_fwd
def varlen_fwd(q, k, v, out_, cu_seqlens_q, cu_seqlens_k, seqused_k, leftpad_k_, block_table_, alibi_slopes_,
               max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_size_left,
               window_size_right, softcap, return_softmax, gen_=None, num_splits=0, seqused_q=None) -> List[torch.Tensor]:
    """mha_varlen_fwd (flash_api.cpp:538-788).  q (total_q,h,d); k/v (total_k,h_k,d) or paged
    (num_blocks,page,h_k,d) with block_table (b,max_blocks) -> [out, softmax_lse(h,total_q), p, rng_state]."""
    lib = _lib.load()
    _common_checks(q, k, v)
    if return_softmax:
        _check(p_dropout > 0.0, "return_softmax is only supported when p_dropout > 0.0")
    _check(cu_seqlens_q.dtype == torch.int32, "cu_seqlens_q must have dtype int32")
    _check(cu_seqlens_k.dtype == torch.int32, "cu_seqlens_k must have dtype int32")
    _check(cu_seqlens_q.is_cuda and cu_seqlens_k.is_cuda, "cu_seqlens must be on CUDA device")
    _check(cu_seqlens_q.is_contiguous() and cu_seqlens_k.is_contiguous(), "cu_seqlens must be contiguous")
    paged = block_table_ is not None
    total_q, h, d = q.shape
    b = cu_seqlens_q.numel() - 1
    _check(b > 0, "batch size must be positive")
    _check(d <= 256, "FlashAttention forward only supports head dimension at most 256")
    _check(d % 8 == 0, "query, key, value, and out_ must have a head_size that is a multiple of 8")
    _check(cu_seqlens_k.numel() == b + 1, "cu_seqlens_k has the wrong shape")
    if paged:
        _check(block_table_.dtype == torch.int32, "block_table must have dtype torch.int32")
        _check(block_table_.stride(-1) == 1, "block_table must have contiguous last dimension")
        num_pages, page, hk = k.shape[0], k.shape[1], k.shape[2]
        _check(page % 256 == 0, "Paged KV cache block size must be divisible by 256")
        _check(num_splits <= 1, "num_splits > 1 is not supported for paged KV in varlen_fwd")
        total_k = 0
    else:
        total_k, hk = k.shape[0], k.shape[1]
        page, num_pages = 0, 0
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    if out_ is not None:
        out = out_
        _check(out.dtype == q.dtype and out.stride(-1) == 1 and tuple(out.shape) == (total_q, h, d), "out has the wrong shape/dtype")
    else:
        out = torch.empty_like(q)
    lse = torch.empty((h, total_q), dtype=torch.float32, device=q.device)
    p = torch.empty(0, dtype=q.dtype, device=q.device)
    a = _lib.FwdArgs()
    if p_dropout > 0.0:
        _check(not paged and seqused_k is None and seqused_q is None,
               "dropout is not supported together with paged KV / seqused")
    rng_state = _dropout_state(a, p_dropout, gen_, q.device, b, h)
    if zero_tensors:
        out.zero_()
        lse.fill_(float("-inf"))
    if total_q == 0:
        return [out, lse, p, rng_state]
    if max_seqlen_k == 0:
        out.zero_()
        lse.fill_(float("inf"))
        return [out, lse, p, rng_state]
    if return_softmax:  # flash_api.cpp:690-694
        p = torch.zeros((b, h, _round128(max_seqlen_q), _round128(max_seqlen_k)), dtype=q.dtype, device=q.device)
        a.s_dmask = _ptr(p)
    a.q, a.k, a.v, a.out, a.softmax_lse = _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse)
    a.q_batch_stride, a.q_row_stride, a.q_head_stride = 0, q.stride(0), q.stride(1)
    a.o_batch_stride, a.o_row_stride, a.o_head_stride = 0, out.stride(0), out.stride(1)
    if paged:
        a.k_batch_stride, a.k_row_stride, a.k_head_stride = k.stride(0), k.stride(1), k.stride(2)
        a.v_batch_stride, a.v_row_stride, a.v_head_stride = v.stride(0), v.stride(1), v.stride(2)
        a.block_table = _ptr(block_table_)
        a.block_table_batch_stride = block_table_.stride(0)
        a.page_block_size, a.num_pages = page, num_pages
    else:
        a.k_batch_stride, a.k_row_stride, a.k_head_stride = 0, k.stride(0), k.stride(1)
        a.v_batch_stride, a.v_row_stride, a.v_head_stride = 0, v.stride(0), v.stride(1)
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, max_seqlen_q, max_seqlen_k, h, hk, d
    a.total_q, a.total_k = total_q, total_k
    a.dtype = _DT[q.dtype]
    a.cu_seqlens_q, a.cu_seqlens_k = _ptr(cu_seqlens_q), _ptr(cu_seqlens_k)
    if seqused_k is not None:
        _check(seqused_k.dtype == torch.int32 and seqused_k.is_cuda and seqused_k.is_contiguous(), "seqused_k must be int32 contiguous cuda")
        a.seqused_k = _ptr(seqused_k)
    if seqused_q is not None:
        _check(seqused_q.dtype == torch.int32 and seqused_q.is_cuda and seqused_q.is_contiguous(), "seqused_q must be int32 contiguous cuda")
        a.seqused_q = _ptr(seqused_q)
    if leftpad_k_ is not None:  # flash_api.cpp:740-747
        _check(not paged, "We don't support Paged KV and leftpad_k running at the same time yet")
        _check(leftpad_k_.dtype == torch.int32, "leftpad_k must have dtype int32")
        _check(leftpad_k_.is_cuda and leftpad_k_.is_contiguous() and tuple(leftpad_k_.shape) == (b,), "leftpad_k must be (batch,) contiguous cuda")
        a.leftpad_k = _ptr(leftpad_k_)
    a.softmax_scale, a.softcap = softmax_scale, softcap
    a.is_causal, a.window_size_left, a.window_size_right = int(is_causal), window_size_left, window_size_right
    a.num_splits = 1
    _alibi(a, alibi_slopes_, b, h)
    with torch.cuda.device(q.device):
        _lib.check(lib.fa_b200_mha_varlen_fwd(C.byref(a), _stream(q)))
    return [out, lse, p, rng_state]


# --------------------------------------------------------------------------------------- fwd_kvcache
def fwd_kvcache(q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_, leftpad_k_,
                block_table_, alibi_slopes_, out_, softmax_scale, is_causal, window_size_left, window_size_right,
                softcap, is_rotary_interleaved, num_splits) -> List[torch.Tensor]:
    """mha_fwd_kvcache (flash_api.cpp:1243-1532).  q (b,sq,h,d); kcache/vcache (b_c,sk,h_k,d) or paged
    (num_blocks,page,h_k,d) -> [out, softmax_lse(b,h,sq)]."""
    lib = _lib.load()
    _common_checks(q, kcache, vcache)
    b, sq, h, d = q.shape
    paged = block_table_ is not None
    if paged:
        _check(block_table_.dtype == torch.int32, "block_table must have dtype torch.int32")
        _check(block_table_.stride(-1) == 1, "block_table must have contiguous last dimension")
        num_pages, page, hk = kcache.shape[0], kcache.shape[1], kcache.shape[2]
        _check(page % 256 == 0, "Paged KV cache block size must be divisible by 256")
        sk = block_table_.shape[1] * page
    else:
        _check(cache_batch_idx_ is not None or kcache.shape[0] == b, "kcache batch must equal q batch unless cache_batch_idx is given")
        sk, hk = kcache.shape[1], kcache.shape[2]
        num_pages, page = 0, 0
    _check(d <= 256, "FlashAttention forward only supports head dimension at most 256")
    _check(d % 8 == 0, "head_size must be a multiple of 8")
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    if seqlens_k_ is not None:
        _check(seqlens_k_.dtype == torch.int32, "seqlens_k must have dtype int32")
        _check(seqlens_k_.is_cuda and seqlens_k_.is_contiguous() and seqlens_k_.numel() == b, "seqlens_k must be (batch,) contiguous cuda")
    for name, t in (("cache_batch_idx", cache_batch_idx_), ("leftpad_k", leftpad_k_)):
        if t is not None:
            _check(t.dtype == torch.int32 and t.is_cuda and t.is_contiguous() and t.numel() == b, f"{name} must be (batch,) int32 contiguous cuda")
            _check(not paged, "We don't support Paged KV and leftpad_k / cache_batch_idx running at the same time yet")
    if leftpad_k_ is not None:
        _check(seqlens_k_ is not None, "leftpad_k requires seqlens_k")
    if rotary_cos_ is not None:
        _check(k_ is not None, "If rotary cos/sin are provided, new key / value to be appended to KV cache must also be provided")
        _check(rotary_sin_ is not None and rotary_cos_.dtype == q.dtype and rotary_sin_.dtype == q.dtype, "rotary_cos/sin must have the same dtype as query")
        _check(rotary_cos_.is_contiguous() and rotary_sin_.is_contiguous(), "rotary_cos/sin must be contiguous")
        _check((rotary_cos_.shape[-1] * 2) % 16 == 0 and rotary_cos_.shape[-1] * 2 <= d, "Only rotary dimensions divisible by 16 are currently supported")
    if k_ is not None:
        # append new keys/values into the cache at position cache_seqlens (flash_fwd_kernel.h:691-790)
        _check(v_ is not None, "If key is supplied, value must also be passed in")
        _check(seqlens_k_ is not None, "If key is supplied, seqlens_k must also be passed in")
        _check(k_.dtype == q.dtype and v_.dtype == q.dtype, "Key and value must have the same dtype as query")
        _check(k_.stride(-1) == 1 and v_.stride(-1) == 1, "Key / value tensor must have contiguous last dimension")
        from .kvcache_ops import append_kv_rotary

        # query i is rotated at cache_seqlens + i when causal or local, else all queries at cache_seqlens
        local = window_size_left >= 0 or window_size_right >= 0
        q = append_kv_rotary(q, kcache, vcache, k_, v_, seqlens_k_, rotary_cos_, rotary_sin_, cache_batch_idx_,
                             leftpad_k_, block_table_, is_rotary_interleaved, bool(is_causal) or local)
        seqlens_eff = seqlens_k_ + k_.shape[1]
    else:
        seqlens_eff = seqlens_k_
    if out_ is not None:
        out = out_
        _check(out.dtype == q.dtype and out.stride(-1) == 1 and tuple(out.shape) == (b, sq, h, d), "out has the wrong shape/dtype")
    else:
        out = torch.empty_like(q)
    lse = torch.empty((b, h, sq), dtype=torch.float32, device=q.device)
    a = _lib.FwdArgs()
    a.q, a.k, a.v, a.out, a.softmax_lse = _ptr(q), _ptr(kcache), _ptr(vcache), _ptr(out), _ptr(lse)
    a.q_batch_stride, a.q_row_stride, a.q_head_stride = q.stride(0), q.stride(1), q.stride(2)
    a.k_batch_stride, a.k_row_stride, a.k_head_stride = kcache.stride(0), kcache.stride(1), kcache.stride(2)
    a.v_batch_stride, a.v_row_stride, a.v_head_stride = vcache.stride(0), vcache.stride(1), vcache.stride(2)
    a.o_batch_stride, a.o_row_stride, a.o_head_stride = out.stride(0), out.stride(1), out.stride(2)
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, sq, sk, h, hk, d
    a.dtype = _DT[q.dtype]
    if paged:
        a.block_table = _ptr(block_table_)
        a.block_table_batch_stride = block_table_.stride(0)
        a.page_block_size, a.num_pages = page, num_pages
    if seqlens_eff is not None:
        seqlens_eff = seqlens_eff.contiguous()
        a.seqused_k = _ptr(seqlens_eff)
    a.cache_batch_idx, a.leftpad_k = _ptr(cache_batch_idx_), _ptr(leftpad_k_)
    a.b_kvcache = 0 if paged else kcache.shape[0]
    a.softmax_scale, a.softcap = softmax_scale, softcap
    a.is_causal, a.window_size_left, a.window_size_right = int(is_causal), window_size_left, window_size_right
    a.num_splits = num_splits
    a.entry = _lib.ENTRY_KVCACHE   # the workspace query must plan like the kvcache launch does
    _alibi(a, alibi_slopes_, b, h)
    with torch.cuda.device(q.device):
        ws = _workspace(lib, a, q.device)  # noqa: F841
        _lib.check(lib.fa_b200_mha_fwd_kvcache(C.byref(a), _stream(q)))
    return [out, lse]


# --------------------------------------------------------------------------------------- bwd
def _bwd_common(a, lib, dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes_, b, h, softmax_scale, is_causal,
                window_size_left, window_size_right, softcap, deterministic):
    a.dout, a.q, a.k, a.v, a.out, a.softmax_lse = _ptr(dout), _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(softmax_lse)
    a.dq, a.dk, a.dv = _ptr(dq), _ptr(dk), _ptr(dv)
    a.dtype = _DT[q.dtype]
    a.softmax_scale, a.softcap = softmax_scale, softcap
    a.is_causal, a.window_size_left, a.window_size_right = int(is_causal), window_size_left, window_size_right
    a.deterministic = int(deterministic)
    if alibi_slopes_ is not None:
        _check(alibi_slopes_.dtype == torch.float32 and alibi_slopes_.is_cuda and alibi_slopes_.stride(-1) == 1, "ALiBi slopes must be fp32 cuda contiguous")
        _check(tuple(alibi_slopes_.shape) in ((h,), (b, h)), "ALiBi slopes must have shape (nheads,) or (batch, nheads)")
        a.alibi_slopes = _ptr(alibi_slopes_)
        a.alibi_slopes_batch_stride = alibi_slopes_.stride(0) if alibi_slopes_.dim() == 2 else 0
    nbytes = lib.fa_b200_bwd_workspace_bytes(C.byref(a))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=q.device)
    a.workspace, a.workspace_bytes = _ptr(ws), nbytes
    return ws


def _alloc_or_check(t_, like, name):
    if t_ is None:
        return torch.empty_like(like)
    _check(t_.dtype == like.dtype, f"{name} must have the same dtype as q")
    _check(t_.is_cuda and t_.stride(-1) == 1, f"{name} must have contiguous last dimension")
    _check(tuple(t_.shape) == tuple(like.shape), f"{name} has the wrong shape")
    return t_


def _bwd_dropout(a, p_dropout, rng_state, softcap):
    """p_dropout + the forward's rng_state (flash_api.cpp:990-1003); the kernel reads (seed, offset) from the device tensor."""
    _check(0.0 <= p_dropout < 1.0, "p_dropout must be in [0, 1)")
    if p_dropout == 0.0:
        return
    _check(softcap == 0.0, "Softcapping does not support dropout for now")
    _check(rng_state is not None, "fa_b200: the backward of a dropout forward needs its rng_state")
    _check(rng_state.is_cuda and rng_state.dtype == torch.int64 and rng_state.numel() == 2 and rng_state.is_contiguous(),
           "rng_state must be a contiguous CUDA int64 tensor of 2 elements")
    a.p_dropout = p_dropout
    a.rng_state = _ptr(rng_state)


def bwd(dout, q, k, v, out, softmax_lse, dq_, dk_, dv_, alibi_slopes_, p_dropout, softmax_scale, is_causal,
        window_size_left, window_size_right, softcap, deterministic, gen_=None, rng_state=None) -> List[torch.Tensor]:
    """mha_bwd (flash_api.cpp:800-1008) -> [dq, dk, dv, softmax_d(b,h,round128(sq))]."""
    lib = _lib.load()
    _common_checks(q, k, v)
    _check(dout.dtype == q.dtype and out.dtype == q.dtype, "query and dout/out must have the same dtype")
    _check(dout.stride(-1) == 1 and out.stride(-1) == 1, "dout/out tensor must have contiguous last dimension")
    b, sq, h, d = q.shape
    sk, hk = k.shape[1], k.shape[2]
    _check(d % 8 == 0, "head_size should be a multiple of 8")
    _check(d <= 256, "FlashAttention backward only supports head dimension at most 256")
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    dq = _alloc_or_check(dq_, q, "dq")
    dk = _alloc_or_check(dk_, k, "dk")
    dv = _alloc_or_check(dv_, v, "dv")
    sq_r = (sq + 127) // 128 * 128
    softmax_d = torch.empty((b, h, sq_r), dtype=torch.float32, device=q.device)
    if sq == 0:
        return [dq, dk, dv, softmax_d]
    if sk == 0:  # flash_api.cpp:995-1000
        dk.zero_(); dv.zero_(); dq.zero_(); softmax_d.zero_()
        return [dq, dk, dv, softmax_d]
    a = _lib.BwdArgs()
    a.softmax_d = _ptr(softmax_d)
    _bwd_dropout(a, p_dropout, rng_state, softcap)
    for name, t in (("q", q), ("k", k), ("v", v), ("o", out), ("do", dout), ("dq", dq), ("dk", dk), ("dv", dv)):
        setattr(a, name + "_batch_stride", t.stride(0))
        setattr(a, name + "_row_stride", t.stride(1))
        setattr(a, name + "_head_stride", t.stride(2))
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, sq, sk, h, hk, d
    with torch.cuda.device(q.device):
        ws = _bwd_common(a, lib, dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes_, b, h, softmax_scale,  # noqa: F841
                         is_causal, window_size_left, window_size_right, softcap, deterministic)
        _lib.check(lib.fa_b200_mha_bwd(C.byref(a), _stream(q)))
    return [dq, dk, dv, softmax_d]


def varlen_bwd(dout, q, k, v, out, softmax_lse, dq_, dk_, dv_, cu_seqlens_q, cu_seqlens_k, alibi_slopes_,
               max_seqlen_q, max_seqlen_k, p_dropout, softmax_scale, zero_tensors, is_causal, window_size_left,
               window_size_right, softcap, deterministic, gen_=None, rng_state=None) -> List[torch.Tensor]:
    """mha_varlen_bwd (flash_api.cpp:1010-1241) -> [dq, dk, dv, softmax_d(h, total_q + 128*b)]."""
    lib = _lib.load()
    _common_checks(q, k, v)
    _check(cu_seqlens_q.dtype == torch.int32 and cu_seqlens_k.dtype == torch.int32, "cu_seqlens must have dtype int32")
    total_q, h, d = q.shape
    total_k, hk = k.shape[0], k.shape[1]
    b = cu_seqlens_q.numel() - 1
    _check(d % 8 == 0, "head_size should be a multiple of 8")
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    dq = _alloc_or_check(dq_, q, "dq")
    dk = _alloc_or_check(dk_, k, "dk")
    dv = _alloc_or_check(dv_, v, "dv")
    softmax_d = torch.empty((h, total_q + 128 * b), dtype=torch.float32, device=q.device)
    if zero_tensors:
        dq.zero_(); dk.zero_(); dv.zero_(); softmax_d.zero_()
    if total_q == 0:
        return [dq, dk, dv, softmax_d]
    if max_seqlen_k == 0 or total_k == 0:
        dk.zero_(); dv.zero_(); dq.zero_(); softmax_d.zero_()
        return [dq, dk, dv, softmax_d]
    a = _lib.BwdArgs()
    a.softmax_d = _ptr(softmax_d)
    _bwd_dropout(a, p_dropout, rng_state, softcap)
    for name, t in (("q", q), ("k", k), ("v", v), ("o", out), ("do", dout), ("dq", dq), ("dk", dk), ("dv", dv)):
        setattr(a, name + "_batch_stride", 0)
        setattr(a, name + "_row_stride", t.stride(0))
        setattr(a, name + "_head_stride", t.stride(1))
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, max_seqlen_q, max_seqlen_k, h, hk, d
    a.total_q, a.total_k = total_q, total_k
    a.cu_seqlens_q, a.cu_seqlens_k = _ptr(cu_seqlens_q), _ptr(cu_seqlens_k)
    with torch.cuda.device(q.device):
        ws = _bwd_common(a, lib, dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes_, b, h, softmax_scale,  # noqa: F841
                         is_causal, window_size_left, window_size_right, softcap, deterministic)
        _lib.check(lib.fa_b200_mha_varlen_bwd(C.byref(a), _stream(q)))
    return [dq, dk, dv, softmax_d]


def fwd_combine(o_partial: torch.Tensor, lse_partial: torch.Tensor, out_dtype=torch.bfloat16):
    """flash_attn_combine (flash_attn/cute/interface.py:3687): o_partial (splits, ..., d) fp32,
    lse_partial (splits, ...) fp32 -> (out (..., d), lse (...))."""
    lib = _lib.load()
    _check(o_partial.dtype == torch.float32 and lse_partial.dtype == torch.float32, "partials must be fp32")
    o_partial, lse_partial = o_partial.contiguous(), lse_partial.contiguous()
    splits, d = o_partial.shape[0], o_partial.shape[-1]
    rows = lse_partial[0].numel()
    out = torch.empty(o_partial.shape[1:], dtype=out_dtype, device=o_partial.device)
    lse = torch.empty(lse_partial.shape[1:], dtype=torch.float32, device=o_partial.device)
    with torch.cuda.device(o_partial.device):
        _lib.check(lib.fa_b200_fwd_combine(_ptr(o_partial), _ptr(lse_partial), _ptr(out), _ptr(lse), splits, rows, d,
                                           _DT[out_dtype], _stream(o_partial)))
    return out, lse

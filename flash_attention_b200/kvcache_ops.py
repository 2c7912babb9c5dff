"""Decode prologue ops on top of the C ABI: in-place KV-cache append (+ rotary on new keys and q)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def append_kv_rotary(q, kcache, vcache, knew, vnew, cache_seqlens, rotary_cos, rotary_sin, cache_batch_idx, leftpad_k,
                     block_table, interleaved, q_pos_per_row):
    """Writes knew/vnew into the cache at rows cache_seqlens[b] + t (rotated when cos/sin are given) and
    returns the (rotated) q to attend with.  In place on kcache / vcache, like the reference."""
    lib = _lib.load()
    a = _lib.AppendArgs()
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    b, sq, h, d = q.shape
    hk = kcache.shape[2]
    a.kcache, a.vcache = p(kcache), p(vcache)
    a.kc_batch_stride, a.kc_row_stride, a.kc_head_stride = kcache.stride(0), kcache.stride(1), kcache.stride(2)
    a.vc_batch_stride, a.vc_row_stride, a.vc_head_stride = vcache.stride(0), vcache.stride(1), vcache.stride(2)
    if knew is not None:
        a.knew, a.vnew = p(knew), p(vnew)
        a.kn_batch_stride, a.kn_row_stride, a.kn_head_stride = knew.stride(0), knew.stride(1), knew.stride(2)
        a.vn_batch_stride, a.vn_row_stride, a.vn_head_stride = vnew.stride(0), vnew.stride(1), vnew.stride(2)
        a.seqlen_new = knew.shape[1]
    q_out = q
    if rotary_cos is not None:
        q_out = torch.empty_like(q)
        a.q, a.q_out = p(q), p(q_out)
        a.q_batch_stride, a.q_row_stride, a.q_head_stride = q.stride(0), q.stride(1), q.stride(2)
        a.qo_batch_stride, a.qo_row_stride, a.qo_head_stride = q_out.stride(0), q_out.stride(1), q_out.stride(2)
        a.rotary_cos, a.rotary_sin = p(rotary_cos), p(rotary_sin)
        a.rotary_dim = rotary_cos.shape[-1] * 2
        a.is_rotary_interleaved = int(interleaved)
        a.q_pos_per_row = int(q_pos_per_row)
    a.cache_seqlens = p(cache_seqlens)
    a.cache_batch_idx, a.leftpad_k = p(cache_batch_idx), p(leftpad_k)
    if block_table is not None:
        a.block_table = p(block_table)
        a.block_table_batch_stride = block_table.stride(0)
        a.page_block_size = kcache.shape[1]
    a.b, a.seqlen_q, a.h, a.h_k, a.d = b, sq, h, hk, d
    a.dtype = _lib.BF16 if q.dtype == torch.bfloat16 else _lib.FP16
    with torch.cuda.device(q.device):
        _lib.check(lib.fa_b200_kvcache_append(C.byref(a), C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)))
    return q_out

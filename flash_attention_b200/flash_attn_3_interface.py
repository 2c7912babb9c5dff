"""FA3-style operator surface on the fa_b200 kernels: the names and argument meaning of the
reference's `hopper/flash_attn_interface.py` (flash_attn_func :809, flash_attn_varlen_func :890,
flash_attn_combine :938, flash_attn_with_kvcache :942, get_scheduler_metadata :1106): fp8 descales on every
entry point, `pack_gqa`, `num_splits`, `seqused_q/k`, `page_table`, public combine.

Arguments this library has no kernel for raise instead of being ignored: `qv`, `attention_chunk`,
`rotary_seqlens`, `cu_seqlens_k_new`.  `sm_margin` is accepted and ignored (grids are not persistent yet).
Differences in return convention follow FA3: `(out, softmax_lse)` when `return_attn_probs=True`.
"""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import flash_attn_2_cuda as _op
from . import flash_attn_interface as _fa2


def _unsupported(**kw):
    for name, (val, default) in kw.items():
        if not (val is None if default is None else val == default):
            raise RuntimeError(f"fa_b200: `{name}` is not supported by this build")


def flash_attn_func(q, k, v, softmax_scale=None, causal=False, qv=None, q_descale=None, k_descale=None, v_descale=None,
                    window_size=(-1, -1), attention_chunk=0, softcap=0.0, num_splits=1, pack_gqa=None,
                    deterministic=False, sm_margin=0, return_attn_probs=False):
    """q (b, sq, h, d), k/v (b, sk, h_k, d); bf16 / fp16 (differentiable) or float8_e4m3fn (forward, bf16 out)."""
    _unsupported(qv=(qv, None), attention_chunk=(attention_chunk, 0))
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    if q.dtype == torch.float8_e4m3fn:
        out, lse = _op.fwd_fp8(q, k, v, None, q_descale, k_descale, v_descale, softmax_scale, causal, window_size[0],
                               window_size[1], softcap, None, num_splits, pack_gqa)
        return (out, lse) if return_attn_probs else out
    if num_splits != 1 or pack_gqa is not None:
        # forward-only fast path with explicit scheduling knobs (no autograd graph, like FA3's inference use)
        if not (q.requires_grad or k.requires_grad or v.requires_grad):
            out, lse = _op.fwd_ex(q, k, v, softmax_scale, causal, window_size[0], window_size[1], softcap, num_splits, pack_gqa)
            return (out, lse) if return_attn_probs else out
    res = _fa2.flash_attn_func(q, k, v, 0.0, softmax_scale, causal, window_size, softcap, None, deterministic,
                               return_attn_probs)
    return (res[0], res[1]) if return_attn_probs else res


def flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, seqused_q=None,
                           seqused_k=None, softmax_scale=None, causal=False, qv=None, q_descale=None, k_descale=None,
                           v_descale=None, window_size=(-1, -1), attention_chunk=0, softcap=0.0, num_splits=1,
                           pack_gqa=None, deterministic=False, sm_margin=0, return_attn_probs=False):
    """Packed variable-length batch: q (total_q, h, d), k/v (total_k, h_k, d)."""
    _unsupported(qv=(qv, None), attention_chunk=(attention_chunk, 0), q_descale=(q_descale, None),
                 k_descale=(k_descale, None), v_descale=(v_descale, None))
    if seqused_q is not None or seqused_k is not None:
        if softmax_scale is None:
            softmax_scale = q.shape[-1] ** (-0.5)
        out, lse, _, _ = _op.varlen_fwd(q, k, v, None, cu_seqlens_q, cu_seqlens_k, seqused_k, None, None, None,
                                        max_seqlen_q, max_seqlen_k, 0.0, softmax_scale, False, causal, window_size[0],
                                        window_size[1], softcap, False, None, 0, seqused_q)
        return (out, lse) if return_attn_probs else out
    res = _fa2.flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, 0.0, softmax_scale,
                                      causal, window_size, softcap, None, deterministic, return_attn_probs)
    return (res[0], res[1]) if return_attn_probs else res


def flash_attn_combine(out_partial, lse_partial, out=None, out_dtype=None):
    """out_partial (splits, ..., d) fp32, lse_partial (splits, ...) fp32 -> (out, lse)."""
    o, lse = _op.fwd_combine(out_partial, lse_partial, out_dtype or torch.bfloat16)
    if out is not None:
        out.copy_(o)
        o = out
    return o, lse


def flash_attn_with_kvcache(q, k_cache, v_cache, k=None, v=None, qv=None, rotary_cos=None, rotary_sin=None,
                            cache_seqlens: Optional[Union[int, torch.Tensor]] = None, cache_batch_idx=None,
                            cache_leftpad=None, page_table=None, cu_seqlens_q=None, cu_seqlens_k_new=None,
                            max_seqlen_q=None, rotary_seqlens=None, q_descale=None, k_descale=None, v_descale=None,
                            softmax_scale=None, causal=False, window_size=(-1, -1), attention_chunk=0, softcap=0.0,
                            rotary_interleaved=True, scheduler_metadata=None, num_splits=0, pack_gqa=None, sm_margin=0,
                            return_softmax_lse=False):
    """Decode / incremental prefill against a (paged) KV cache; `page_table` is FA2's `block_table`."""
    _unsupported(qv=(qv, None), cu_seqlens_q=(cu_seqlens_q, None), cu_seqlens_k_new=(cu_seqlens_k_new, None),
                 rotary_seqlens=(rotary_seqlens, None), attention_chunk=(attention_chunk, 0), q_descale=(q_descale, None),
                 k_descale=(k_descale, None), v_descale=(v_descale, None))
    if scheduler_metadata is not None and num_splits == 0:
        num_splits = int(scheduler_metadata[0])
    return _fa2.flash_attn_with_kvcache(q, k_cache, v_cache, k, v, rotary_cos, rotary_sin, cache_seqlens, cache_batch_idx,
                                        cache_leftpad, page_table, softmax_scale, causal, window_size, softcap,
                                        rotary_interleaved, None, num_splits, return_softmax_lse)


def get_scheduler_metadata(batch_size, max_seqlen_q, max_seqlen_k, num_heads_q, num_heads_kv, headdim, cache_seqlens,
                           qkv_dtype=torch.bfloat16, headdim_v=None, cu_seqlens_q=None, cu_seqlens_k_new=None,
                           cache_leftpad=None, page_size=None, max_seqlen_k_new=0, causal=False, window_size=(-1, -1),
                           attention_chunk=0, has_softcap=False, num_splits=0, pack_gqa=None, sm_margin=0):
    """The plan the library would pick for this decode shape, computed once and re-usable across steps
    (FA3: hopper/flash_prepare_scheduler.cu).  Returns an int32 tensor [num_splits, gqa_pack_factor, num_m_blocks]."""
    splits, gp, mblocks = _op.plan_kvcache(batch_size, max_seqlen_q, max_seqlen_k, num_heads_q, num_heads_kv, headdim,
                                           page_size or 0, causal, window_size[0], window_size[1], num_splits, pack_gqa)
    return torch.tensor([splits, gp, mblocks], dtype=torch.int32, device=cache_seqlens.device)

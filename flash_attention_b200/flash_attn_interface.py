"""Public operator API, mirroring flash_attn/flash_attn_interface.py of the reference
(flash_attn_func :1156, flash_attn_varlen_func :1391, flash_attn_with_kvcache :1485 and the packed
variants :1019-1388): same names, argument meaning, return values and error behaviour, implemented
on top of this package's `flash_attn_2_cuda` operator module."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import flash_attn_2_cuda as _op
from . import ops as _ops  # noqa: F401  registers the fa_b200::{fwd,varlen_fwd,bwd,varlen_bwd} custom ops + fakes


def _tracing() -> bool:
    """True while torch.compile / export captures a graph: route through the registered custom ops (opaque to the tracer,
    shapes from their fakes); eager calls go straight to the compiled module."""
    return torch.compiler.is_compiling()


def _c(x):
    return x if x is None or x.stride(-1) == 1 else x.contiguous()


def _pad8(t, d):
    return t if d % 8 == 0 else torch.nn.functional.pad(t, [0, 8 - d % 8])


class _FlashAttnFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic,
                return_softmax):
        if softmax_scale is None:
            softmax_scale = q.shape[-1] ** (-0.5)
        d = q.shape[-1]
        qp, kp, vp = (_c(_pad8(t, d)) for t in (q, k, v))
        if _tracing():
            out_p, lse, S_dmask, rng_state = torch.ops.fa_b200.fwd(qp, kp, vp, dropout_p, softmax_scale, causal, window_size[0],
                                                                   window_size[1], softcap, alibi_slopes,
                                                                   return_softmax and dropout_p > 0)
        else:
            out_p, lse, S_dmask, rng_state = _op.fwd(qp, kp, vp, None, alibi_slopes, dropout_p, softmax_scale, causal,
                                                     window_size[0], window_size[1], softcap,
                                                     return_softmax and dropout_p > 0, None)
        ctx.save_for_backward(qp, kp, vp, out_p, lse, rng_state)
        ctx.args = (dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic, d)
        out = out_p[..., :d]
        return (out, lse, S_dmask) if return_softmax else out

    @staticmethod
    def backward(ctx, dout, *unused):
        qp, kp, vp, out_p, lse, rng_state = ctx.saved_tensors
        dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic, d = ctx.args
        dq, dk, dv = torch.empty_like(qp), torch.empty_like(kp), torch.empty_like(vp)
        dout_p = _c(_pad8(dout, d))
        if _tracing():
            torch.ops.fa_b200.bwd(dout_p, qp, kp, vp, out_p, lse, dq, dk, dv, dropout_p, softmax_scale, causal, window_size[0],
                                  window_size[1], softcap, alibi_slopes, deterministic, rng_state)
        else:
            _op.bwd(dout_p, qp, kp, vp, out_p, lse, dq, dk, dv, alibi_slopes, dropout_p, softmax_scale, causal,
                    window_size[0], window_size[1], softcap, deterministic, None, rng_state)
        return dq[..., :d], dk[..., :d], dv[..., :d], None, None, None, None, None, None, None, None


class _FlashAttnVarlenFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_q, cu_k, max_q, max_k, dropout_p, softmax_scale, causal, window_size, softcap,
                alibi_slopes, deterministic, return_softmax, block_table):
        if softmax_scale is None:
            softmax_scale = q.shape[-1] ** (-0.5)
        d = q.shape[-1]
        qp, kp, vp = (_c(_pad8(t, d)) for t in (q, k, v))
        if _tracing():
            out_p, lse, S_dmask, rng_state = torch.ops.fa_b200.varlen_fwd(
                qp, kp, vp, cu_q, cu_k, max_q, max_k, dropout_p, softmax_scale, causal, window_size[0], window_size[1], softcap,
                alibi_slopes, return_softmax and dropout_p > 0, block_table, None, None, False)
        else:
            out_p, lse, S_dmask, rng_state = _op.varlen_fwd(qp, kp, vp, None, cu_q, cu_k, None, None, block_table,
                                                            alibi_slopes, max_q, max_k, dropout_p, softmax_scale, False,
                                                            causal, window_size[0], window_size[1], softcap,
                                                            return_softmax and dropout_p > 0, None)
        ctx.save_for_backward(qp, kp, vp, out_p, lse, cu_q, cu_k, rng_state)
        ctx.args = (max_q, max_k, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic, d)
        out = out_p[..., :d]
        return (out, lse, S_dmask) if return_softmax else out

    @staticmethod
    def backward(ctx, dout, *unused):
        qp, kp, vp, out_p, lse, cu_q, cu_k, rng_state = ctx.saved_tensors
        max_q, max_k, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic, d = ctx.args
        dq, dk, dv = torch.empty_like(qp), torch.empty_like(kp), torch.empty_like(vp)
        dout_p = _c(_pad8(dout, d))
        if _tracing():
            torch.ops.fa_b200.varlen_bwd(dout_p, qp, kp, vp, out_p, lse, dq, dk, dv, cu_q, cu_k, max_q, max_k, dropout_p,
                                         softmax_scale, causal, window_size[0], window_size[1], softcap, alibi_slopes,
                                         deterministic, rng_state, False)
        else:
            _op.varlen_bwd(dout_p, qp, kp, vp, out_p, lse, dq, dk, dv, cu_q, cu_k, alibi_slopes, max_q, max_k, dropout_p,
                           softmax_scale, False, causal, window_size[0], window_size[1], softcap, deterministic, None,
                           rng_state)
        return (dq[..., :d], dk[..., :d], dv[..., :d]) + (None,) * 13


def flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                    alibi_slopes=None, deterministic=False, return_attn_probs=False, q_descale=None, k_descale=None,
                    v_descale=None):
    """q (b, sq, h, d), k/v (b, sk, h_k, d) -> out (b, sq, h, d) [, softmax_lse (b, h, sq), S_dmask].
    Causal / sliding-window masks are aligned to the bottom-right corner when sq != sk.
    float8_e4m3fn inputs (forward only, FA3-style `*_descale` (b, h_k) fp32 factors) return bf16."""
    if q.dtype == torch.float8_e4m3fn:
        assert dropout_p == 0.0, "fp8 forward has no dropout"
        if softmax_scale is None:
            softmax_scale = q.shape[-1] ** (-0.5)
        out, lse = _op.fwd_fp8(_c(q), _c(k), _c(v), None, q_descale, k_descale, v_descale, softmax_scale, causal,
                               window_size[0], window_size[1], softcap, alibi_slopes)
        return (out, lse, None) if return_attn_probs else out
    if not _tracing() and not (torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)):
        # inference fast path: no autograd node, straight into the compiled module
        d = q.shape[-1]
        if softmax_scale is None:
            softmax_scale = d ** (-0.5)
        qp, kp, vp = (_c(_pad8(t, d)) for t in (q, k, v))
        out_p, lse, S_dmask, _ = _op.fwd(qp, kp, vp, None, alibi_slopes, dropout_p, softmax_scale, causal, window_size[0],
                                         window_size[1], softcap, return_attn_probs and dropout_p > 0, None)
        out = out_p if d % 8 == 0 else out_p[..., :d]
        return (out, lse, S_dmask) if return_attn_probs else out
    return _FlashAttnFunc.apply(q, k, v, dropout_p, softmax_scale, causal, tuple(window_size), softcap, alibi_slopes,
                                deterministic, return_attn_probs)


def flash_attn_qkvpacked_func(qkv, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                              alibi_slopes=None, deterministic=False, return_attn_probs=False):
    """qkv (b, s, 3, h, d)."""
    return flash_attn_func(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], dropout_p, softmax_scale, causal, window_size,
                           softcap, alibi_slopes, deterministic, return_attn_probs)


def flash_attn_kvpacked_func(q, kv, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                             alibi_slopes=None, deterministic=False, return_attn_probs=False):
    """q (b, sq, h, d), kv (b, sk, 2, h_k, d)."""
    return flash_attn_func(q, kv[:, :, 0], kv[:, :, 1], dropout_p, softmax_scale, causal, window_size, softcap,
                           alibi_slopes, deterministic, return_attn_probs)


def flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p=0.0,
                           softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0, alibi_slopes=None,
                           deterministic=False, return_attn_probs=False, block_table=None):
    """q (total_q, h, d), k/v (total_k, h_k, d) [or paged (num_blocks, page, h_k, d) with block_table]."""
    return _FlashAttnVarlenFunc.apply(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p,
                                      softmax_scale, causal, tuple(window_size), softcap, alibi_slopes, deterministic,
                                      return_attn_probs, block_table)


def flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False,
                                     window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False,
                                     return_attn_probs=False):
    """qkv (total, 3, h, d)."""
    return flash_attn_varlen_func(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens, cu_seqlens, max_seqlen, max_seqlen,
                                  dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic,
                                  return_attn_probs)


def flash_attn_varlen_kvpacked_func(q, kv, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p=0.0,
                                    softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                                    alibi_slopes=None, deterministic=False, return_attn_probs=False):
    """q (total_q, h, d), kv (total_k, 2, h_k, d)."""
    return flash_attn_varlen_func(q, kv[:, 0], kv[:, 1], cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                  dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic,
                                  return_attn_probs)


def flash_attn_with_kvcache(q, k_cache, v_cache, k=None, v=None, rotary_cos=None, rotary_sin=None, cache_seqlens=None,
                            cache_batch_idx=None, cache_leftpad=None, block_table=None, softmax_scale=None,
                            causal=False, window_size=(-1, -1), softcap=0.0, rotary_interleaved=True,
                            alibi_slopes=None, num_splits=0, return_softmax_lse=False):
    """Decode / incremental prefill against a (possibly paged) KV cache; appends k/v in place when given."""
    assert k_cache.stride(-1) == 1 and v_cache.stride(-1) == 1, "k_cache/v_cache must have contiguous last dimension"
    q, k, v = _c(q), _c(k), _c(v)
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    if cache_seqlens is not None and isinstance(cache_seqlens, int):
        cache_seqlens = torch.full((k_cache.shape[0],), cache_seqlens, dtype=torch.int32, device=k_cache.device)
    out, lse = _op.fwd_kvcache(q, k_cache, v_cache, k, v, cache_seqlens, rotary_cos, rotary_sin, cache_batch_idx,
                               cache_leftpad, block_table, alibi_slopes, None, softmax_scale, causal, window_size[0],
                               window_size[1], softcap, rotary_interleaved, num_splits)
    return (out, lse) if return_softmax_lse else out


def flash_attn_combine(out_partial, lse_partial, out_dtype=torch.bfloat16):
    """Combine split attention partials: out_partial (splits, ..., d) fp32, lse_partial (splits, ...) fp32."""
    return _op.fwd_combine(out_partial, lse_partial, out_dtype)

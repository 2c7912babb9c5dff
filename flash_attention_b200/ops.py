"""torch.compile-visible operator layer of this package: `torch.library.custom_op`s in the `fa_b200` namespace with fake
(meta) implementations, mirroring the reference's `flash_attn::_flash_attn_forward / _varlen_forward / _backward /
_varlen_backward` (flash_attn/flash_attn_interface.py:84-144, 153-243, 252-338, 347-452).

The real implementations call the compiled operator module; the fakes state the output shapes a tracer needs:
  forward   (out like q, softmax_lse (b, h, sq) | varlen (h, total_q) fp32, S_dmask (b, h, round128(sq), round128(sk)) or
             empty, rng_state (2,) int64)
  backward  mutates dq / dk / dv in place and returns softmax_d (b, h, round128(sq)) | varlen (h, total_q + 128 * b) fp32.
`flash_attn_interface.py` routes through these ops while a graph is being captured (torch.compile / export) and calls the
compiled module directly in eager mode (the custom-op dispatcher costs more host time than the launch itself).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import flash_attn_2_cuda as _op


def _round128(x: int) -> int:
    return (x + 127) // 128 * 128


# --------------------------------------------------------------------------------------------------------- forward
@torch.library.custom_op("fa_b200::fwd", mutates_args=(), device_types="cuda")
def fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dropout_p: float, softmax_scale: float, causal: bool,
        window_size_left: int, window_size_right: int, softcap: float, alibi_slopes: Optional[torch.Tensor],
        return_softmax: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    out, lse, p, rng_state = _op.fwd(q, k, v, None, alibi_slopes, dropout_p, softmax_scale, causal, window_size_left,
                                     window_size_right, softcap, return_softmax, None)
    return out, lse, p, rng_state


@fwd.register_fake
def _(q, k, v, dropout_p, softmax_scale, causal, window_size_left, window_size_right, softcap, alibi_slopes, return_softmax):
    b, sq, h, _ = q.shape
    sk = k.shape[1]
    out = torch.empty_like(q)
    lse = q.new_empty((b, h, sq), dtype=torch.float32)
    p = q.new_empty((b, h, _round128(sq), _round128(sk)) if return_softmax else (0,))
    return out, lse, p, q.new_empty((2,), dtype=torch.int64)


@torch.library.custom_op("fa_b200::varlen_fwd", mutates_args=(), device_types="cuda")
def varlen_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens_q: torch.Tensor, cu_seqlens_k: torch.Tensor,
               max_seqlen_q: int, max_seqlen_k: int, dropout_p: float, softmax_scale: float, causal: bool,
               window_size_left: int, window_size_right: int, softcap: float, alibi_slopes: Optional[torch.Tensor],
               return_softmax: bool, block_table: Optional[torch.Tensor], leftpad_k: Optional[torch.Tensor],
               seqused_k: Optional[torch.Tensor], zero_tensors: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    out, lse, p, rng_state = _op.varlen_fwd(q, k, v, None, cu_seqlens_q, cu_seqlens_k, seqused_k, leftpad_k, block_table,
                                            alibi_slopes, max_seqlen_q, max_seqlen_k, dropout_p, softmax_scale, zero_tensors,
                                            causal, window_size_left, window_size_right, softcap, return_softmax, None)
    return out, lse, p, rng_state


@varlen_fwd.register_fake
def _(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p, softmax_scale, causal, window_size_left,
      window_size_right, softcap, alibi_slopes, return_softmax, block_table, leftpad_k, seqused_k, zero_tensors):
    total_q, h, _ = q.shape
    b = cu_seqlens_q.numel() - 1
    out = torch.empty_like(q)
    lse = q.new_empty((h, total_q), dtype=torch.float32)
    p = q.new_empty((b, h, _round128(max_seqlen_q), _round128(max_seqlen_k)) if return_softmax else (0,))
    return out, lse, p, q.new_empty((2,), dtype=torch.int64)


# --------------------------------------------------------------------------------------------------------- backward
@torch.library.custom_op("fa_b200::bwd", mutates_args=("dq", "dk", "dv"), device_types="cuda")
def bwd(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, softmax_lse: torch.Tensor,
        dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, dropout_p: float, softmax_scale: float, causal: bool,
        window_size_left: int, window_size_right: int, softcap: float, alibi_slopes: Optional[torch.Tensor],
        deterministic: bool, rng_state: Optional[torch.Tensor]) -> torch.Tensor:
    return _op.bwd(dout, q, k, v, out, softmax_lse, dq, dk, dv, alibi_slopes, dropout_p, softmax_scale, causal,
                   window_size_left, window_size_right, softcap, deterministic, None, rng_state)[3]


@bwd.register_fake
def _(dout, q, k, v, out, softmax_lse, dq, dk, dv, dropout_p, softmax_scale, causal, window_size_left, window_size_right,
      softcap, alibi_slopes, deterministic, rng_state):
    b, sq, h, _ = q.shape
    return q.new_empty((b, h, _round128(sq)), dtype=torch.float32)


@torch.library.custom_op("fa_b200::varlen_bwd", mutates_args=("dq", "dk", "dv"), device_types="cuda")
def varlen_bwd(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor,
               softmax_lse: torch.Tensor, dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, cu_seqlens_q: torch.Tensor,
               cu_seqlens_k: torch.Tensor, max_seqlen_q: int, max_seqlen_k: int, dropout_p: float, softmax_scale: float,
               causal: bool, window_size_left: int, window_size_right: int, softcap: float,
               alibi_slopes: Optional[torch.Tensor], deterministic: bool, rng_state: Optional[torch.Tensor],
               zero_tensors: bool) -> torch.Tensor:
    return _op.varlen_bwd(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k, alibi_slopes, max_seqlen_q,
                          max_seqlen_k, dropout_p, softmax_scale, zero_tensors, causal, window_size_left, window_size_right,
                          softcap, deterministic, None, rng_state)[3]


@varlen_bwd.register_fake
def _(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p,
      softmax_scale, causal, window_size_left, window_size_right, softcap, alibi_slopes, deterministic, rng_state, zero_tensors):
    total_q, h, _ = q.shape
    b = cu_seqlens_q.numel() - 1
    return q.new_empty((h, total_q + 128 * b), dtype=torch.float32)

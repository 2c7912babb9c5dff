"""CPU oracle for the fused-attention hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch fp32 on the host CPU, the math that the
reference's attention kernels implement.  It is the checker for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``flash_attention_b200/``
imports it and the product path raises if the CUDA library is missing.

Parity pinning: ``oracle/make_golden.py`` runs the reference's own
``attention_ref`` (extracted from ``/root/reference/tests/test_flash_attn.py:217-304``
in the build container) on seeded inputs and stores the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
those fixtures.  => parity is PINNED.

What is restated (reference file:line):
  * score/mask/softmax/PV pipeline        tests/test_flash_attn.py:217-304
  * bottom-right aligned causal / window  tests/test_flash_attn.py:182-214,
                                          csrc/flash_attn/src/mask.h:172-173
  * ALiBi bias                            tests/test_flash_attn.py:29-55,
                                          csrc/flash_attn/src/mask.h:148-165
  * tanh soft-cap                         csrc/flash_attn/flash_api.cpp:130-134
  * LSE definition (natural log)          csrc/flash_attn/src/softmax.h:169-186
  * GQA head mapping h -> h // (H/H_k)    csrc/flash_attn/src/flash_fwd_kernel.h:155
  * fp8 forward (descales, bf16 out, e4m3 `intermediate_dtype`)   flash_attn/cute/testing.py:326-470
                                          (pinned by tests/golden/fp8_*.npz, oracle/make_golden.py)
  * split-KV combine                      csrc/flash_attn/src/flash_fwd_kernel.h:1117-1299
  * paged-KV gather                       csrc/flash_attn/src/flash_fwd_kernel.h:592-597
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

NEG_INF = float("-inf")


def visible_mask(
    seqlen_q: int,
    seqlen_k: int,
    causal: bool = False,
    window: Tuple[int, int] = (-1, -1),
    q_lens: Optional[torch.Tensor] = None,  # (b,) valid query count per batch
    k_lens: Optional[torch.Tensor] = None,  # (b,) valid key count per batch
    device=None,
) -> Optional[torch.Tensor]:
    """Boolean (b|1, 1, sq, sk) tensor, True where key j is visible to query i.

    Key j is visible to query i iff  lo(i) <= j <= hi(i)  and j < k_len, with
      shift = k_len - q_len                      (bottom-right alignment)
      hi(i) = i + shift + window_right           (window_right = 0 when causal)
      lo(i) = i + shift - window_left            (only if window_left >= 0)
    """
    wl, wr = window
    if causal:
        wr = 0
    nb = 1
    if q_lens is not None:
        nb = q_lens.numel()
    if k_lens is not None:
        nb = k_lens.numel()
    i = torch.arange(seqlen_q, device=device).view(1, 1, seqlen_q, 1)
    j = torch.arange(seqlen_k, device=device).view(1, 1, 1, seqlen_k)
    sk = (
        torch.full((nb, 1, 1, 1), seqlen_k, device=device, dtype=torch.long)
        if k_lens is None
        else k_lens.view(nb, 1, 1, 1).long()
    )
    sq = (
        torch.full((nb, 1, 1, 1), seqlen_q, device=device, dtype=torch.long)
        if q_lens is None
        else q_lens.view(nb, 1, 1, 1).long()
    )
    vis = j < sk
    if wl < 0 and wr < 0:
        return vis if k_lens is not None else None
    shift = sk - sq
    if wr >= 0:
        vis = vis & (j <= i + shift + wr)
    if wl >= 0:
        vis = vis & (j >= i + shift - wl)
    return vis


def alibi_bias(
    slopes: torch.Tensor,  # (h,) or (b,h) fp32
    seqlen_q: int,
    seqlen_k: int,
    q_lens: Optional[torch.Tensor] = None,
    k_lens: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """bias[b,h,i,j] = -slope[b,h] * |i + (k_len - q_len) - j|.

    (For causal masks the reference adds +slope*j instead, which differs from this
    by a per-row constant and therefore yields the same softmax; mask.h:148-165.)
    """
    if slopes.dim() == 1:
        slopes = slopes.unsqueeze(0)
    b = slopes.shape[0]
    dev = slopes.device
    i = torch.arange(seqlen_q, device=dev).view(1, 1, seqlen_q, 1)
    j = torch.arange(seqlen_k, device=dev).view(1, 1, 1, seqlen_k)
    sk = seqlen_k if k_lens is None else k_lens.view(-1, 1, 1, 1).long()
    sq = seqlen_q if q_lens is None else q_lens.view(-1, 1, 1, 1).long()
    rel = (i + sk - sq - j).abs().to(torch.float32)
    return -slopes.view(b, -1, 1, 1).to(torch.float32) * rel


def attention_oracle(
    q: torch.Tensor,  # (b, sq, h, d)
    k: torch.Tensor,  # (b, sk, h_k, d)
    v: torch.Tensor,  # (b, sk, h_k, dv)
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    window: Tuple[int, int] = (-1, -1),
    softcap: float = 0.0,
    alibi_slopes: Optional[torch.Tensor] = None,
    q_lens: Optional[torch.Tensor] = None,
    k_lens: Optional[torch.Tensor] = None,
    compute_dtype: torch.dtype = torch.float32,
    p_dtype: Optional[torch.dtype] = None,  # quantise P like the kernel does (bf16 / e4m3)
    q_descale: Optional[torch.Tensor] = None,  # (b, h_k) fp8 only
    k_descale: Optional[torch.Tensor] = None,
    v_descale: Optional[torch.Tensor] = None,
    lse_masked_value: float = float("inf"),
    dropout_p: float = 0.0,
    dropout_mask: Optional[torch.Tensor] = None,  # (b, h, sq, sk) bool, True = keep (tests/test_flash_attn.py:231,288-296)
    p_quant: str = "unnormalized",  # what `p_dtype` rounds: "unnormalized" = exp(s - m) as the kernels do;
                                    # "normalized" = softmax(s), the reference's `intermediate_dtype`
                                    # (flash_attn/cute/testing.py:461-462)
    reorder_ops: bool = False,      # scale K before the score GEMM and keep the scores in `compute_dtype`
                                    # (the reference's low-precision baseline, flash_attn/cute/testing.py:381-386)
):
    """Returns (out (b,sq,h,dv) in q's dtype [bf16 if q is fp8], lse (b,h,sq) fp32).

    Rows with no visible key give out = 0 and lse = +inf (FA2 non-split convention,
    softmax.h:180) unless ``lse_masked_value`` says otherwise (-inf for split partials).
    """
    out_dtype = q.dtype
    if q.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
        out_dtype = torch.bfloat16
    b, sq, h, d = q.shape
    sk, hk = k.shape[1], k.shape[2]
    g = h // hk
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(d)
    qf, kf, vf = q.to(compute_dtype), k.to(compute_dtype), v.to(compute_dtype)
    # de-scales are applied in fp32 and rounded back to the compute dtype (flash_attn/cute/testing.py:363-369)
    if q_descale is not None:
        qf = (qf.float() * q_descale.float().repeat_interleave(g, dim=1).view(b, 1, h, 1)).to(compute_dtype)
    if k_descale is not None:
        kf = (kf.float() * k_descale.float().view(b, 1, hk, 1)).to(compute_dtype)
    if v_descale is not None:
        vf = (vf.float() * v_descale.float().view(b, 1, hk, 1)).to(compute_dtype)
    if g > 1:
        kf = kf.repeat_interleave(g, dim=2)
        vf = vf.repeat_interleave(g, dim=2)
    if reorder_ops:
        s = torch.einsum("bihd,bjhd->bhij", qf, kf * softmax_scale).float()
    else:
        s = torch.einsum("bihd,bjhd->bhij", qf, kf).float() * softmax_scale
    if softcap > 0:
        s = torch.tanh(s / softcap) * softcap
    if alibi_slopes is not None:
        s = s + alibi_bias(alibi_slopes, sq, sk, q_lens, k_lens)
    vis = visible_mask(sq, sk, causal, window, q_lens, k_lens, device=q.device)
    if vis is not None:
        s = s.masked_fill(~vis, NEG_INF)
    m = s.amax(dim=-1, keepdim=True)
    m_safe = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    e = torch.exp(s - m_safe)
    denom = e.sum(dim=-1, keepdim=True)
    lse = (m_safe + torch.log(denom)).squeeze(-1)
    dead = denom.squeeze(-1) == 0
    lse = torch.where(dead, torch.full_like(lse, lse_masked_value), lse)
    if p_dtype is not None and p_quant == "normalized":
        p = (e / denom.clamp_min(1e-30)).to(compute_dtype).to(p_dtype).to(compute_dtype)
        o = torch.einsum("bhij,bjhd->bihd", p, vf)
    elif p_dtype is not None:
        # the kernels quantise the un-normalised probabilities, then divide
        p = e.to(p_dtype).to(compute_dtype)
        o = torch.einsum("bhij,bjhd->bihd", p, vf) / denom.clamp_min(1e-30).permute(0, 2, 1, 3).to(compute_dtype)
    else:
        p = (e / denom.clamp_min(1e-30)).to(compute_dtype)
        if dropout_mask is not None:  # attention_ref: masked_fill(~dropout_mask, 0) * 1/(1-p)
            p = p.masked_fill(~dropout_mask, 0.0) * (1.0 / (1.0 - dropout_p))
        o = torch.einsum("bhij,bjhd->bihd", p, vf)
    o = o.masked_fill(dead.permute(0, 2, 1).unsqueeze(-1), 0.0)
    if q_lens is not None:
        qvalid = torch.arange(sq, device=q.device).view(1, sq) < q_lens.view(-1, 1)
        o = o.masked_fill(~qvalid.view(b, sq, 1, 1), 0.0)
    return o.to(out_dtype), lse


_PHILOX_M0, _PHILOX_M1, _PHILOX_W0, _PHILOX_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_7(c0, c1, c2, c3, k0, k1, rounds: int = 7):
    """Philox4x32 with 7 rounds (the reference's generator, csrc/flash_attn/src/philox.cuh:29-51; Salmon et al.
    SC'11) on numpy uint64 arrays holding 32-bit values.  Returns the four output words."""
    import numpy as np

    m32 = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & m32 for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    for _ in range(rounds):
        p0 = np.uint64(_PHILOX_M0) * c0
        p1 = np.uint64(_PHILOX_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m32, lo1, (hi0 ^ c3 ^ k1) & m32, lo0
        k0 = (k0 + np.uint64(_PHILOX_W0)) & m32
        k1 = (k1 + np.uint64(_PHILOX_W1)) & m32
    return c0, c1, c2, c3


def philox_keep_mask(seed: int, offset: int, p_dropout: float, b: int, h: int, sq: int, sk: int) -> torch.Tensor:
    """The dropout mask of the reference's kernels, restated on the CPU: (b, h, sq, sk) bool, True = keep.

    FA2 draws one Philox4x32-7 value per thread and 16 x 32 (query x key) block of its mma.sync accumulator
    (csrc/flash_attn/src/dropout.h:20-47, philox.cuh:29-51); per element (query R, key C) of head bh = batch * h + head:
        counter = (lo, hi of [offset + bh * 32 + lane], R // 16, C // 32),  key = (seed_lo, seed_hi)
        lane    = (R % 8) * 4 + (C % 8) // 2
        byte    = ((C % 32) // 16) * 8 + ((C % 16) // 8) * 4 + (C % 2) + 2 * ((R % 16) // 8)   of the 16 output bytes
    keep iff byte <= floor((1 - p_dropout) * 255) (8-bit threshold, dropout.h:17).  The library's kernels use the same
    stream (flash_attention_b200/csrc/fa_dropout.cuh); on the GPU box the mask is also compared with the sign pattern of
    the S_dmask the reference's own FA2 binary returns for the same (seed, offset) (tests/test_dropout_gpu.py)."""
    import numpy as np

    seed &= 2**64 - 1
    offset &= 2**64 - 1
    thr = int(np.floor(np.float32(1.0 - p_dropout) * np.float32(255.0)))
    R = np.arange(sq, dtype=np.uint64).reshape(1, -1, 1)
    C = np.arange(sk, dtype=np.uint64).reshape(1, 1, -1)
    bh = np.arange(b * h, dtype=np.uint64).reshape(-1, 1, 1)
    shape = (b * h, sq, sk)
    lane = (R % 8) * 4 + (C % 8) // 2
    off = (np.uint64(offset) + bh * np.uint64(32) + lane) & np.uint64(2**64 - 1)
    off = np.broadcast_to(off, shape)
    words = philox4x32_7(off & np.uint64(0xFFFFFFFF), off >> np.uint64(32), np.broadcast_to(R // 16, shape),
                         np.broadcast_to(C // 32, shape), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    idx = np.broadcast_to(((C % 32) // 16) * 8 + ((C % 16) // 8) * 4 + (C % 2) + 2 * ((R % 16) // 8), shape).astype(np.int64)
    w = np.stack(words, axis=-1)                                    # (bh, sq, sk, 4) 32-bit words
    word = np.take_along_axis(w, (idx // 4)[..., None], axis=-1)[..., 0]
    byte = (word >> (np.uint64(8) * (idx % 4).astype(np.uint64))) & np.uint64(0xFF)
    keep = byte <= np.uint64(thr)
    return torch.from_numpy(np.ascontiguousarray(keep)).view(b, h, sq, sk)


def attention_oracle_bwd(q, k, v, dout, **kw):
    """(out, lse, dq, dk, dv) with fp32 autograd through ``attention_oracle``.

    Reference for dQ/dK/dV: flash_bwd_kernel.h:457-724 computes exactly the vjp of the
    forward; autograd of the fp32 restatement is the oracle (as in
    tests/test_flash_attn.py:1060-1100)."""
    qf = q.detach().float().requires_grad_(True)
    kf = k.detach().float().requires_grad_(True)
    vf = v.detach().float().requires_grad_(True)
    out, lse = attention_oracle(qf, kf, vf, **kw)
    dq, dk, dv = torch.autograd.grad(out, (qf, kf, vf), dout.float())
    return out.detach(), lse.detach(), dq, dk, dv


def combine_oracle(o_partial: torch.Tensor, lse_partial: torch.Tensor, out_dtype=torch.bfloat16):
    """Split-KV combine.  o_partial (splits, b, h, sq, d) fp32, lse_partial (splits, b, h, sq).

    lse = logsumexp_s(lse_s); out = sum_s exp(lse_s - lse) * o_s.  If every partial is
    -inf the final lse is +inf and out = 0 (flash_fwd_kernel.h:1196,1204)."""
    lse_max = lse_partial.amax(dim=0)
    lse_max_safe = torch.where(torch.isfinite(lse_max), lse_max, torch.zeros_like(lse_max))
    w = torch.exp(lse_partial - lse_max_safe)
    tot = w.sum(dim=0)
    lse = torch.log(tot) + lse_max_safe
    dead = tot == 0
    lse = torch.where(dead, torch.full_like(lse, float("inf")), lse)
    scale = torch.where(dead.unsqueeze(0), torch.zeros_like(w), w / tot.clamp_min(1e-30))
    out = (scale.unsqueeze(-1) * o_partial).sum(dim=0)
    return out.to(out_dtype), lse


def gather_paged_kv(cache: torch.Tensor, block_table: torch.Tensor, seqlen: int) -> torch.Tensor:
    """cache (num_blocks, page, h_k, d), block_table (b, max_blocks) -> (b, seqlen, h_k, d).

    logical position p of batch b lives at cache[block_table[b, p // page], p % page]."""
    page = cache.shape[1]
    nblk = (seqlen + page - 1) // page
    g = cache[block_table[:, :nblk].long()]  # (b, nblk, page, h_k, d)
    b = block_table.shape[0]
    return g.reshape(b, nblk * page, cache.shape[2], cache.shape[3])[:, :seqlen]


def attention_oracle_varlen(q, k, v, cu_q, cu_k, max_q, max_k, **kw):
    """Packed varlen wrapper: q (total_q,h,d), k/v (total_k,h_k,d) -> out (total_q,h,d), lse (h,total_q)."""
    nb = cu_q.numel() - 1
    h, d = q.shape[1], q.shape[2]
    out = torch.zeros_like(q if q.dtype not in (torch.float8_e4m3fn,) else q.to(torch.bfloat16))
    lse = torch.zeros(h, q.shape[0], dtype=torch.float32, device=q.device)
    for i in range(nb):
        a, b_ = int(cu_q[i]), int(cu_q[i + 1])
        c, e = int(cu_k[i]), int(cu_k[i + 1])
        if b_ == a:
            continue
        kw_i = dict(kw)
        if kw_i.get("alibi_slopes") is not None and kw_i["alibi_slopes"].dim() == 2:
            kw_i["alibi_slopes"] = kw_i["alibi_slopes"][i : i + 1]
        if e == c:
            lse[:, a:b_] = float("inf")
            continue
        o, l = attention_oracle(q[a:b_].unsqueeze(0), k[c:e].unsqueeze(0), v[c:e].unsqueeze(0), **kw_i)
        out[a:b_] = o[0]
        lse[:, a:b_] = l[0]
    return out, lse


def attention_flops(b, h, sq, sk, d, causal=False, window=(-1, -1), mode="fwd"):
    """Reference FLOP convention (flash_attn/cute/bench_utils.py:15-48,
    benchmarks/benchmark_flash_attention.py:27-30): fwd = 4*b*h*pairs*d where pairs is
    the number of visible (i,j) pairs; bwd = 2.5x, fwd_bwd = 3.5x."""
    wl, wr = window
    if causal:
        wr = 0
    if wl < 0 and wr < 0:
        pairs = sq * sk
    else:
        i = torch.arange(sq, dtype=torch.long)
        hi = torch.clamp(i + sk - sq + (wr if wr >= 0 else sk), max=sk - 1)
        lo = torch.clamp(i + sk - sq - wl, min=0) if wl >= 0 else torch.zeros_like(i)
        pairs = int(torch.clamp(hi - lo + 1, min=0).sum())
    f = 4 * b * h * pairs * d
    return {"fwd": f, "bwd": 2.5 * f, "fwd_bwd": 3.5 * f}[mode]

"""GPU parity tests of the decode path: flash_attn_with_kvcache with (paged) caches, split-KV + combine,
in-place append of new K/V, rotary, cache_batch_idx, cache_leftpad  (reference: test_flash_attn_kvcache,
tests/test_flash_attn.py:1859-2140; tolerance mult=3 +1e-5 there, the oracle rule here)."""
import math

import pytest
import torch

from oracle.attention_oracle import attention_oracle, gather_paged_kv  # checker only
from tests.util import check_fwd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m
    from flash_attention_b200 import _lib

    _lib.load()
    return m


def rotary_ref(x, cos, sin, pos, interleaved):
    """x (b, s, h, d); pos (b, s) integer positions; cos/sin (seqlen_ro, rd/2).  Plain torch fp32."""
    rd = cos.shape[-1] * 2
    c = cos.float()[pos].unsqueeze(2)  # (b, s, 1, rd/2)
    s_ = sin.float()[pos].unsqueeze(2)
    xr, rest = x.float()[..., :rd], x.float()[..., rd:]
    if interleaved:
        x1, x2 = xr[..., 0::2], xr[..., 1::2]
        o = torch.stack((x1 * c - x2 * s_, x1 * s_ + x2 * c), dim=-1).flatten(-2)
    else:
        x1, x2 = xr[..., : rd // 2], xr[..., rd // 2:]
        o = torch.cat((x1 * c - x2 * s_, x1 * s_ + x2 * c), dim=-1)
    return torch.cat((o, rest), dim=-1).to(x.dtype)


@pytest.mark.parametrize("num_splits", [1, 0, 3])
@pytest.mark.parametrize("paged", [False, True])
def test_decode_cache_seqlens(fab, paged, num_splits):
    torch.manual_seed(0)
    b, h, hk, d, page = 3, 8, 2, 128, 256
    seqlens = torch.tensor([700, 256, 1500], dtype=torch.int32, device="cuda")
    nblk = 6
    q = torch.randn(b, 1, h, d, device="cuda", dtype=torch.bfloat16)
    if paged:
        kc = torch.randn(b * nblk, page, hk, d, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(b * nblk, page, hk, d, device="cuda", dtype=torch.bfloat16)
        bt = torch.randperm(b * nblk, device="cuda").int().view(b, nblk)
        kd, vd = gather_paged_kv(kc.cpu(), bt.cpu(), nblk * page), gather_paged_kv(vc.cpu(), bt.cpu(), nblk * page)
        out, lse = fab.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=seqlens, block_table=bt, num_splits=num_splits,
                                               return_softmax_lse=True)
    else:
        kc = torch.randn(b, nblk * page, hk, d, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(b, nblk * page, hk, d, device="cuda", dtype=torch.bfloat16)
        for i in range(b):  # garbage beyond each batch's valid length must not leak (0 * NaN)
            kc[i, int(seqlens[i]):] = float("nan")
            vc[i, int(seqlens[i]):] = float("nan")
        kd, vd = kc.cpu(), vc.cpu()
        out, lse = fab.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=seqlens, num_splits=num_splits, return_softmax_lse=True)
    kd, vd = torch.nan_to_num(kd), torch.nan_to_num(vd)
    check_fwd(out, lse, q, kd, vd, k_lens=seqlens.cpu())


@pytest.mark.parametrize("interleaved", [False, True])
@pytest.mark.parametrize("paged", [False, True])
@pytest.mark.parametrize("sq", [1, 5])
def test_append_kv_with_rotary(fab, sq, paged, interleaved):
    torch.manual_seed(1)
    b, h, hk, d, page, nblk, rd = 2, 4, 2, 128, 256, 3, 64
    seqlens = torch.tensor([300, 511], dtype=torch.int32, device="cuda")
    q = torch.randn(b, sq, h, d, device="cuda", dtype=torch.bfloat16)
    kn = torch.randn(b, sq, hk, d, device="cuda", dtype=torch.bfloat16)
    vn = torch.randn(b, sq, hk, d, device="cuda", dtype=torch.bfloat16)
    ang = torch.rand(nblk * page, rd // 2, device="cuda") * 2 * math.pi
    cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
    if paged:
        kc = torch.randn(b * nblk, page, hk, d, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(b * nblk, page, hk, d, device="cuda", dtype=torch.bfloat16)
        bt = torch.randperm(b * nblk, device="cuda").int().view(b, nblk)
        k_before = gather_paged_kv(kc.cpu(), bt.cpu(), nblk * page)
        v_before = gather_paged_kv(vc.cpu(), bt.cpu(), nblk * page)
    else:
        kc = torch.randn(b, nblk * page, hk, d, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(b, nblk * page, hk, d, device="cuda", dtype=torch.bfloat16)
        bt = None
        k_before, v_before = kc.cpu().clone(), vc.cpu().clone()
    out = fab.flash_attn_with_kvcache(q, kc, vc, k=kn, v=vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=seqlens,
                                      block_table=bt, causal=True, rotary_interleaved=interleaved)
    # reference: rotate, append, attend
    pos = seqlens.cpu().long().view(b, 1) + torch.arange(sq).view(1, sq)
    k_rot = rotary_ref(kn.cpu(), cos.cpu(), sin.cpu(), pos, interleaved)
    q_rot = rotary_ref(q.cpu(), cos.cpu(), sin.cpu(), pos, interleaved)
    k_ref, v_ref = k_before.clone(), v_before.clone()
    for i in range(b):
        L = int(seqlens[i])
        k_ref[i, L:L + sq] = k_rot[i]
        v_ref[i, L:L + sq] = vn[i].cpu()
    k_after = gather_paged_kv(kc.cpu(), bt.cpu(), nblk * page) if paged else kc.cpu()
    v_after = gather_paged_kv(vc.cpu(), bt.cpu(), nblk * page) if paged else vc.cpu()
    assert torch.equal(v_after, v_ref)                                      # reference: torch.equal(v_cache)
    assert torch.allclose(k_after.float(), k_ref.float(), atol=1.6e-2, rtol=1e-2)  # reference: allclose(k_cache, 1e-3) in fp16
    check_fwd(out, None, q_rot, k_ref, v_ref, causal=True, q_lens=None, k_lens=(seqlens.cpu() + sq))


def test_cache_batch_idx_and_leftpad(fab):
    torch.manual_seed(2)
    b, bc, h, hk, d, sk = 2, 5, 4, 4, 128, 512
    q = torch.randn(b, 1, h, d, device="cuda", dtype=torch.bfloat16)
    kc = torch.randn(bc, sk, hk, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(bc, sk, hk, d, device="cuda", dtype=torch.bfloat16)
    idx = torch.tensor([3, 1], dtype=torch.int32, device="cuda")
    seqlens = torch.tensor([400, 129], dtype=torch.int32, device="cuda")
    leftpad = torch.tensor([17, 128], dtype=torch.int32, device="cuda")
    out = fab.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=seqlens, cache_batch_idx=idx, cache_leftpad=leftpad)
    for i in range(b):
        lo, hi = int(leftpad[i]), int(seqlens[i])
        ks, vs = kc[int(idx[i]), lo:hi].unsqueeze(0), vc[int(idx[i]), lo:hi].unsqueeze(0)
        check_fwd(out[i:i + 1], None, q[i:i + 1], ks, vs)


def test_combine_op_matches_oracle(fab):
    from oracle.attention_oracle import combine_oracle

    torch.manual_seed(3)
    splits, rows, d = 5, 77, 128
    op = torch.randn(splits, rows, d, device="cuda")
    lp = torch.randn(splits, rows, device="cuda") * 3
    lp[:, 5] = float("-inf")          # a row every split left empty
    lp[2, 9] = float("-inf")
    out, lse = fab.flash_attn_combine(op, lp, out_dtype=torch.bfloat16)
    ro, rl = combine_oracle(op.cpu(), lp.cpu())
    fin = torch.isfinite(rl)
    assert torch.allclose(lse.cpu()[fin], rl[fin], atol=1e-5, rtol=1e-5)   # tests/cute/test_flash_attn_combine.py:45-49
    assert torch.equal(torch.isinf(lse.cpu()), ~fin) and torch.all(out.cpu()[5] == 0)
    assert (out.float().cpu() - ro.float()).abs().max().item() <= 2e-2


def test_paged_block_size_must_be_256_multiple(fab):
    q = torch.randn(1, 1, 2, 128, device="cuda", dtype=torch.bfloat16)
    kc = torch.randn(4, 64, 2, 128, device="cuda", dtype=torch.bfloat16)
    bt = torch.arange(4, device="cuda", dtype=torch.int32).view(1, 4)
    with pytest.raises(RuntimeError, match="divisible by 256"):
        fab.flash_attn_with_kvcache(q, kc, kc, cache_seqlens=torch.tensor([100], dtype=torch.int32, device="cuda"), block_table=bt)

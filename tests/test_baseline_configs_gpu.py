"""BASELINE.json configs at FULL size (C3, C4, C5): the oracle cannot evaluate whole problems of this size on
the host, so each is checked on an oracle-sized slice plus a size-independent property."""
import pytest
import torch

from oracle.attention_oracle import attention_oracle, attention_oracle_bwd, gather_paged_kv  # checker only
from tests.util import check_fwd, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m
    from flash_attention_b200 import _lib

    _lib.load()
    return m


def test_c3_gqa_causal_s16384_fwd_bwd(fab):
    """C3: bf16 fwd+bwd causal GQA 32:4 s16384 d128 (b=2).  Slice: the FIRST 256 queries only see the first 256
    keys (causal, sq == sk) -> out / dQ of that corner vs the oracle; dK/dV are checked through linearity in dO."""
    b, h, hk, s, d = 2, 32, 4, 16384, 128
    q, k, v = make_qkv(b, s, s, h, hk, d, seed=3)
    do = torch.randn_like(q)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    out = fab.flash_attn_func(q, k, v, causal=True)
    out.backward(do)
    bi, kvh, n = 1, 2, 256
    heads = slice(kvh * 8, kvh * 8 + 8)
    sl = lambda t, hs: t[bi:bi + 1, :n, hs].detach().cpu()
    ro, _, rdq, _, _ = attention_oracle_bwd(sl(q, heads), sl(k, slice(kvh, kvh + 1)), sl(v, slice(kvh, kvh + 1)), sl(do, heads), causal=True)
    assert (out[bi:bi + 1, :n, heads].float().cpu() - ro.float()).abs().max().item() <= 2e-2
    assert (q.grad[bi:bi + 1, :n, heads].float().cpu() - rdq).abs().max().item() <= 3e-2 * max(1.0, rdq.abs().max().item())
    assert torch.isfinite(k.grad).all() and torch.isfinite(v.grad).all()
    # property: dV rows are linear in dO (scale dO by 2 on one head group -> dV doubles) checked on the device
    q2, k2, v2 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    fab.flash_attn_func(q2, k2, v2, causal=True).backward(2 * do)
    rel = (v2.grad.float() - 2 * v.grad.float()).abs().max().item() / v.grad.float().abs().max().item()
    assert rel < 2e-2


def test_c4_fp8_window_s32768(fab):
    """C4: fp8 e4m3 fwd s32768 d128 window (4096, 0).  Slice: 256 consecutive queries at the end of the sequence
    see at most the last 4352 keys -> compare with the oracle on that window."""
    f8 = torch.float8_e4m3fn
    b, h, s, d, w = 1, 16, 32768, 128, 4096
    q, k, v = make_qkv(b, s, s, h, h, d, seed=4)
    q8, k8, v8 = q.to(f8), k.to(f8), v.to(f8)
    out = fab.flash_attn_func(q8, k8, v8, causal=True, window_size=(w, 0))
    hi, n = 5, 256
    qs = q8[:, -n:, hi:hi + 1].cpu().float()
    ks = k8[:, -(n + w):, hi:hi + 1].cpu().float()
    vs = v8[:, -(n + w):, hi:hi + 1].cpu().float()
    ref, _ = attention_oracle(qs, ks, vs, causal=True, window=(w, 0))   # bottom-right aligned: same geometry
    refq, _ = attention_oracle(qs, ks, vs, causal=True, window=(w, 0), p_dtype=f8)
    err = (out[:, -n:, hi:hi + 1].float().cpu() - ref).abs().max().item()
    assert err <= 4 * (refq - ref).abs().max().item() + 2e-2, err


def test_c5_paged_decode_kv65536(fab):
    """C5: decode b64 sq1 h32/4 d128 kv65536 paged(256), split-KV auto.  Slice: three batch rows vs the oracle on
    the gathered cache; property: the result does not depend on num_splits (1 vs auto vs 7) beyond fp32 rounding."""
    b, h, hk, d, page, kv = 64, 32, 4, 128, 256, 65536
    nb = kv // page
    kc = torch.randn(b * nb, page, hk, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(b * nb, page, hk, d, device="cuda", dtype=torch.bfloat16)
    bt = torch.randperm(b * nb, device="cuda").int().view(b, nb)
    cs = torch.full((b,), kv, dtype=torch.int32, device="cuda")
    cs[7] = 40000
    q = torch.randn(b, 1, h, d, device="cuda", dtype=torch.bfloat16)
    outs = {ns: fab.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cs, block_table=bt, num_splits=ns) for ns in (0, 1, 7)}
    assert (outs[0].float() - outs[1].float()).abs().max().item() < 2e-3
    assert (outs[7].float() - outs[1].float()).abs().max().item() < 2e-3
    for bi in (0, 7, 63):
        kd = gather_paged_kv(kc.cpu(), bt[bi:bi + 1].cpu(), kv)
        vd = gather_paged_kv(vc.cpu(), bt[bi:bi + 1].cpu(), kv)
        check_fwd(outs[0][bi:bi + 1], None, q[bi:bi + 1], kd, vd, k_lens=cs[bi:bi + 1].cpu())

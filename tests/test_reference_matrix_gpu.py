"""A seeded sample of the reference's own test matrix (tests/test_flash_attn.py: test_flash_attn_qkvpacked :566-700,
test_flash_attn_output :861-1140): head dims 32..256 incl. non-multiples of 8, its sequence-length pairs, MHA / MQA / GQA,
causal / local / ALiBi, dropout 0.17, deterministic, both dtypes -- checked the way the reference checks itself:
|out - ref_fp32| <= 2 |ref_lowp - ref_fp32| + atol and 3x for dq/dk/dv, dropout mask taken from the sign of S_dmask, dropout
fraction within 0.01 (:1121-1132).  Head dims above 128 are forward-only in this build: the backward must raise."""
import random

import pytest
import torch

from oracle.attention_oracle import attention_oracle  # noqa: F401  (checker only, through tests.util)
from tests.util import bwd_tolerances, fwd_tolerances

pytestmark = pytest.mark.gpu

DS = [32, 40, 59, 64, 80, 96, 111, 128, 160, 192, 224, 256]
SEQS = [(113, 203), (128, 217), (113, 211), (108, 256), (256, 512), (512, 256), (1024, 1024), (1023, 1024), (1024, 1023), (97, 97),
        (200, 200), (384, 384), (768, 768), (1025, 1025)]


def _cases(n=44, seed=0):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        d = DS[i % len(DS)]
        sq, sk = rng.choice(SEQS)
        mha = rng.choice(["mha", "mqa", "gqa"])
        causal, local = rng.random() < 0.5, rng.random() < 0.35
        alibi = rng.random() < 0.3
        p = rng.choice([0.0, 0.0, 0.17])
        det = rng.random() < 0.4
        dtype = rng.choice([torch.float16, torch.bfloat16])
        out.append(pytest.param(i, d, sq, sk, mha, causal, local, alibi, p, det, dtype,
                                id=f"{i}-d{d}-s{sq}x{sk}-{mha}{'-causal' if causal else ''}{'-local' if local else ''}"
                                   f"{'-alibi' if alibi else ''}-p{p}{'-det' if det else ''}-{str(dtype)[6:]}"))
    return out


@pytest.mark.parametrize("i,d,sq,sk,mha,causal,local,alibi,p,det,dtype", _cases())
def test_reference_matrix_sample(i, d, sq, sk, mha, causal, local, alibi, p, det, dtype):
    import flash_attention_b200 as fab

    torch.manual_seed(i)
    b, h = 2, 6
    hk = {"mha": 6, "mqa": 1, "gqa": 2}[mha]
    g = torch.Generator().manual_seed(i)
    window = (-1, -1) if not local else tuple(int(x) for x in torch.randint(0, sk, (2,), generator=g))
    q = torch.randn(b, sq, h, d, generator=g).to(dtype).cuda().requires_grad_(True)
    k = torch.randn(b, sk, hk, d, generator=g).to(dtype).cuda().requires_grad_(True)
    v = torch.randn(b, sk, hk, d, generator=g).to(dtype).cuda().requires_grad_(True)
    slopes = (torch.rand(b, h, generator=g) * 0.3).float().cuda() if alibi else None
    torch.cuda.manual_seed(i)
    out, lse, s_dmask = fab.flash_attn_func(q, k, v, p, causal=causal, window_size=window, alibi_slopes=slopes, deterministic=det,
                                            return_attn_probs=True)
    kw = dict(causal=causal, window=window, alibi_slopes=slopes)
    if p > 0:
        keep = (s_dmask[:, :, :sq, :sk] >= 0).cpu()
        kw.update(dropout_p=p, dropout_mask=keep)
    ref32, lse_ref, bound = fwd_tolerances(q.detach(), k.detach(), v.detach(), **kw)
    err = (out.detach().float().cpu() - ref32).abs().max().item()
    assert err <= bound, f"out {err:.3e} > {bound:.3e}"
    if p > 0:  # dropout fraction among the visible entries (tests/test_flash_attn.py:1125-1127)
        from oracle.attention_oracle import visible_mask

        vis = visible_mask(sq, sk, causal, window, None, None, device="cpu")
        valid = keep if vis is None else keep[vis.expand_as(keep)]
        frac = 1.0 - valid.float().mean().item()
        assert abs(frac - p) < 0.015, frac
    do = torch.randn(b, sq, h, d, generator=g).to(dtype).cuda()
    out.backward(do)
    tol = bwd_tolerances(q.detach(), k.detach(), v.detach(), do, **kw)
    for name, gt in (("dq", q.grad), ("dk", k.grad), ("dv", v.grad)):
        ref, bnd = tol[name]
        e = (gt.float().cpu() - ref).abs().max().item()
        assert not torch.isnan(gt).any() and e <= bnd, f"{name}: {e:.3e} > {bnd:.3e}"

"""A seeded sample of the reference's varlen test matrix (tests/test_flash_attn.py:1143-1460, test_flash_attn_varlen_output):
random per-sequence lengths (its "full" / "random" / "third" padding modes), head dims 32..256, MHA / MQA / GQA, causal / local,
ALiBi, dropout 0.17, deterministic -- forward (2x rule) and backward (3x rule) per sequence against the oracle; head dims above
128 are forward-only in this build."""
import random

import pytest
import torch

from tests.util import bwd_tolerances, fwd_tolerances

pytestmark = pytest.mark.gpu

DS = [32, 59, 64, 80, 96, 128, 160, 256]
PAIRS = [(1, 147), (113, 203), (128, 217), (113, 211), (108, 256), (256, 512), (512, 256), (1024, 1024), (1023, 1024), (1024, 1023)]


def _lens(rng, b, smax, mode):
    if mode == "full":
        return [smax] * b
    if mode == "random":
        return [rng.randint(max(1, smax - 20), smax) for _ in range(b)]
    return [rng.randint(max(1, smax // 3), smax) for _ in range(b)]  # "third"


def _cases(n=32, seed=2):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        d = DS[i % len(DS)]
        sq, sk = rng.choice(PAIRS)
        mha = rng.choice(["mha", "mqa", "gqa"])
        causal, local = rng.random() < 0.5, rng.random() < 0.3
        alibi = rng.random() < 0.25
        p = rng.choice([0.0, 0.0, 0.17])
        det = rng.random() < 0.3
        mode = rng.choice(["full", "random", "third"])
        out.append(pytest.param(i, d, sq, sk, mha, causal, local, alibi, p, det, mode,
                                id=f"{i}-d{d}-s{sq}x{sk}-{mha}-{mode}{'-causal' if causal else ''}{'-local' if local else ''}"
                                   f"{'-alibi' if alibi else ''}-p{p}{'-det' if det else ''}"))
    return out


@pytest.mark.parametrize("i,d,sq,sk,mha,causal,local,alibi,p,det,mode", _cases())
def test_varlen_matrix_sample(i, d, sq, sk, mha, causal, local, alibi, p, det, mode):
    import flash_attention_b200 as fab

    rng = random.Random(1000 + i)
    g = torch.Generator().manual_seed(1000 + i)
    dt = torch.bfloat16 if i % 2 else torch.float16
    b, h = 3, 6
    hk = {"mha": 6, "mqa": 1, "gqa": 2}[mha]
    lq, lk = _lens(rng, b, sq, mode), _lens(rng, b, sk, mode)
    cu_q = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), dtype=torch.int32).cuda()
    cu_k = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), dtype=torch.int32).cuda()
    q = torch.randn(sum(lq), h, d, generator=g).to(dt).cuda().requires_grad_(True)
    k = torch.randn(sum(lk), hk, d, generator=g).to(dt).cuda().requires_grad_(True)
    v = torch.randn(sum(lk), hk, d, generator=g).to(dt).cuda().requires_grad_(True)
    do = torch.randn(sum(lq), h, d, generator=g).to(dt).cuda()
    window = (-1, -1) if not local else tuple(int(x) for x in torch.randint(0, min(lk), (2,), generator=g))
    slopes = (torch.rand(b, h, generator=g) * 0.3).float().cuda() if alibi else None
    torch.cuda.manual_seed(i)
    out, lse, s_dmask = fab.flash_attn_varlen_func(q, k, v, cu_q, cu_k, max(lq), max(lk), p, causal=causal, window_size=window,
                                                   alibi_slopes=slopes, deterministic=det, return_attn_probs=True)
    out.backward(do)
    for j in range(b):
        qs, ks = slice(int(cu_q[j]), int(cu_q[j + 1])), slice(int(cu_k[j]), int(cu_k[j + 1]))
        kw = dict(causal=causal, window=window, alibi_slopes=slopes[j:j + 1] if slopes is not None else None)
        if p > 0:
            kw.update(dropout_p=p, dropout_mask=(s_dmask[j:j + 1, :, :lq[j], :lk[j]] >= 0).cpu())
        qj, kj, vj = q[qs][None].detach(), k[ks][None].detach(), v[ks][None].detach()
        ref32, lse_ref, bound = fwd_tolerances(qj, kj, vj, **kw)
        err = (out[qs][None].detach().float().cpu() - ref32).abs().max().item()
        assert err <= bound, f"seq {j} out {err:.3e} > {bound:.3e}"
        fin = torch.isfinite(lse_ref)
        lj = lse[:, qs].cpu()[None]
        assert torch.equal(torch.isfinite(lj), fin)
        if fin.any():
            assert (lj[fin] - lse_ref[fin]).abs().max().item() <= 2e-3 * max(1.0, lse_ref[fin].abs().max().item())
        tol = bwd_tolerances(qj, kj, vj, do[qs][None], **kw)
        for name, gt in (("dq", q.grad[qs]), ("dk", k.grad[ks]), ("dv", v.grad[ks])):
            ref, bnd = tol[name]
            e = (gt[None].float().cpu() - ref).abs().max().item()
            assert not torch.isnan(gt).any() and e <= bnd, f"seq {j} {name}: {e:.3e} > {bnd:.3e}"

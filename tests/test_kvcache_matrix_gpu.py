"""A seeded sample of the reference's kv-cache test matrix (tests/test_flash_attn.py:1859-2140, test_flash_attn_kvcache):
combinations of paged / dense cache, in-place append (+ rotary, interleaved or not, full or half rotary dim), split-KV,
MHA / MQA / GQA, causal / local, ALiBi, cache_batch_idx, cache_leftpad, its (seqlen_q, seqlen_k) pairs and head dims.
Each batch element is checked on its own against the oracle over its valid key range (the reference compares with
mult = 3 to its pytorch reference; here the 2x rule of tests/util.py), and the appended cache contents are checked like the
reference does (v: torch.equal, k: allclose)."""
import math
import random

import pytest
import torch

from oracle.attention_oracle import gather_paged_kv  # checker only
from tests.test_kvcache_gpu import rotary_ref
from tests.util import check_fwd

pytestmark = pytest.mark.gpu

PAIRS = [(1, 128), (1, 339), (3, 1024), (64, 800), (64, 256), (3, 799), (64, 2048), (16, 5000), (1, 16384), (128, 128)]


def _cases(n=36, seed=1):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        sq, sk = rng.choice(PAIRS)
        d = rng.choice([64, 128, 128, 128, 256])
        paged = rng.random() < 0.5
        new_kv = rng.random() < 0.5
        rot = rng.choice([0.0, 0.5, 1.0]) if new_kv else 0.0
        inter = rng.random() < 0.5
        mha = rng.choice(["mha", "mqa", "gqa"])
        causal, local = rng.random() < 0.5, rng.random() < 0.3
        alibi = rng.random() < 0.25 and rot == 0.0
        extra = "none" if paged or rot > 0 else rng.choice(["none", "batch_idx", "leftpad", "both"])
        ns = rng.choice([0, 1, 1, 3])
        out.append(pytest.param(i, sq, sk, d, paged, new_kv, rot, inter, mha, causal, local, alibi, extra, ns,
                                id=f"{i}-s{sq}x{sk}-d{d}-{'paged' if paged else 'dense'}{'-append' if new_kv else ''}"
                                   f"{'-rot%.1f' % rot if rot else ''}-{mha}{'-causal' if causal else ''}{'-local' if local else ''}"
                                   f"{'-alibi' if alibi else ''}-{extra}-ns{ns}"))
    return out


@pytest.mark.parametrize("i,sq,sk,d,paged,new_kv,rot,inter,mha,causal,local,alibi,extra,ns", _cases())
def test_kvcache_matrix_sample(i, sq, sk, d, paged, new_kv, rot, inter, mha, causal, local, alibi, extra, ns):
    import flash_attention_b200 as fab

    g = torch.Generator().manual_seed(100 + i)
    dt = torch.bfloat16
    b, h = 2, 6
    hk = {"mha": 6, "mqa": 1, "gqa": 3}[mha]
    page = 256
    s_new = sq if new_kv else 0
    sk_r = (sk + page - 1) // page * page
    nblk = sk_r // page
    bc = b + 2 if extra in ("batch_idx", "both") else b
    rnd = lambda *s: torch.randn(*s, generator=g).to(dt).cuda()
    q = rnd(b, sq, h, d)
    kn, vn = (rnd(b, sq, hk, d), rnd(b, sq, hk, d)) if new_kv else (None, None)
    window = (-1, -1) if not local else tuple(int(x) for x in torch.randint(0, sk, (2,), generator=g))
    leftpad = torch.tensor([int(torch.randint(0, max(1, sk // 4), (1,), generator=g)) for _ in range(b)], dtype=torch.int32) \
        if extra in ("leftpad", "both") else None
    lo = [int(leftpad[j]) if leftpad is not None else 0 for j in range(b)]
    seqlens = torch.tensor([int(torch.randint(lo[j] + 1, sk - s_new + 1, (1,), generator=g)) if sk - s_new > lo[j] + 1 else sk - s_new
                            for j in range(b)], dtype=torch.int32)
    idx = torch.randperm(bc, generator=g)[:b].int() if bc != b else None
    slopes = (torch.rand(b, h, generator=g) * 0.3).float().cuda() if alibi else None
    if paged:
        kc, vc = rnd(b * nblk, page, hk, d), rnd(b * nblk, page, hk, d)
        bt = torch.randperm(b * nblk, generator=g).int().view(b, nblk).cuda()
        k_before, v_before = gather_paged_kv(kc.cpu(), bt.cpu(), sk_r), gather_paged_kv(vc.cpu(), bt.cpu(), sk_r)
    else:
        kc, vc = rnd(bc, sk_r, hk, d), rnd(bc, sk_r, hk, d)
        bt = None
        k_before, v_before = kc.cpu().clone(), vc.cpu().clone()
    cos = sin = None
    rd = int(rot * d) // 16 * 16
    if rd > 0:
        ang = torch.rand(sk_r + sq, rd // 2, generator=g) * 2 * math.pi
        cos, sin = ang.cos().to(dt).cuda(), ang.sin().to(dt).cuda()
    res = fab.flash_attn_with_kvcache(q, kc, vc, k=kn, v=vn, rotary_cos=cos, rotary_sin=sin, cache_seqlens=seqlens.cuda(),
                                      cache_batch_idx=idx.cuda() if idx is not None else None,
                                      cache_leftpad=leftpad.cuda() if leftpad is not None else None, block_table=bt, causal=causal,
                                      window_size=window, rotary_interleaved=inter, alibi_slopes=slopes, num_splits=ns,
                                      return_softmax_lse=True)
    out, lse = res
    k_after = gather_paged_kv(kc.cpu(), bt.cpu(), sk_r) if paged else kc.cpu()
    v_after = gather_paged_kv(vc.cpu(), bt.cpu(), sk_r) if paged else vc.cpu()
    for j in range(b):
        row = int(idx[j]) if idx is not None else j
        L = int(seqlens[j])
        qj = q[j:j + 1].cpu()
        kj, vj = k_before[row:row + 1].clone(), v_before[row:row + 1].clone()
        if new_kv:
            knj = kn[j:j + 1].cpu()
            if rd > 0:
                pos_k = (L + torch.arange(sq)).view(1, sq)
                pos_q = pos_k if (causal or local) else torch.full((1, sq), L)
                knj = rotary_ref(knj, cos.cpu(), sin.cpu(), pos_k, inter)
                qj = rotary_ref(qj, cos.cpu(), sin.cpu(), pos_q, inter)
            kj[0, L:L + sq] = knj[0]
            vj[0, L:L + sq] = vn[j].cpu()
            assert torch.equal(v_after[row, L:L + sq], vj[0, L:L + sq])
            assert torch.allclose(k_after[row, L:L + sq].float(), kj[0, L:L + sq].float(), atol=1.6e-2, rtol=1e-2)
        Lt = L + s_new
        ks, vs = kj[:, lo[j]:Lt], vj[:, lo[j]:Lt]
        check_fwd(out[j:j + 1], lse[j:j + 1], qj, ks, vs, name=f"batch {j}", causal=causal, window=window,
                  alibi_slopes=slopes[j:j + 1].cpu() if slopes is not None else None)

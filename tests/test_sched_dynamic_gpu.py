"""Dynamic work-item scheduling of the persistent kernels (FwdSched / Bwd2Sched): with `sm_margin` the grid is shrunk to a
few CTAs, so that small problems have many more items than CTAs and every item after the first is drawn from the global
counter.  Checked against the oracle, and bit-for-bit against the full-grid launch (one item per CTA, static) where the
arithmetic per item is the same (forward out / lse, dK / dV, deterministic dQ).  (The reference's analogue: the
DynamicPersistentTileScheduler runs under the same tests as the static ones, hopper/test_flash_attn.py.)"""
import pytest
import torch

from tests.util import bwd_tolerances, check_fwd, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_dynamic(monkeypatch):
    # the forward only draws dynamically from 24 key blocks per item on (shorter items: static deal); 2 = always
    monkeypatch.setenv("FA_B200_SCHED_DYNAMIC", "2")


@pytest.fixture(scope="module")
def ops():
    import flash_attention_b200 as m

    m.flash_attn_2_cuda  # loads the operator module, which registers torch.ops.flash_attn_3
    return torch.ops.flash_attn_3


def _margin(ctas):
    return torch.cuda.get_device_properties(0).multi_processor_count - ctas


CASES = [
    # b, sq, sk, h, hk, d, causal, window
    (2, 1024, 1024, 4, 4, 128, False, (-1, -1)),
    (2, 1024, 1024, 4, 4, 128, True, (-1, -1)),
    (1, 1500, 2000, 6, 2, 128, True, (-1, -1)),
    (2, 777, 1333, 4, 1, 64, False, (300, 100)),
    (1, 2048, 2048, 2, 2, 256, True, (-1, -1)),
]


@pytest.mark.parametrize("b,sq,sk,h,hk,d,causal,window", CASES)
@pytest.mark.parametrize("ctas", [2, 7])
def test_forward_few_ctas(ops, b, sq, sk, h, hk, d, causal, window, ctas):
    q, k, v = make_qkv(b, sq, sk, h, hk, d)
    kw = dict(is_causal=causal, window_size_left=window[0], window_size_right=window[1])
    out0, lse0, _, _ = ops.fwd(q, k, v, **kw)
    out1, lse1, _, _ = ops.fwd(q, k, v, sm_margin=_margin(ctas), **kw)
    assert torch.equal(out0, out1) and torch.equal(lse0, lse1)
    check_fwd(out1, lse1, q, k, v, causal=causal, window=window)
    import os
    os.environ["FA_B200_SCHED_DYNAMIC"] = "0"   # and the static deal of the same few-CTA grid
    out2, lse2, _, _ = ops.fwd(q, k, v, sm_margin=_margin(ctas), **kw)
    assert torch.equal(out0, out2) and torch.equal(lse0, lse2)


def test_forward_varlen_few_ctas(ops):
    torch.manual_seed(1)
    lens = [700, 1, 129, 2048, 300, 64]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    tot, h, hk, d = sum(lens), 4, 2, 128
    q = torch.randn(tot, h, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(tot, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(tot, hk, d, device="cuda", dtype=torch.bfloat16)
    kw = dict(cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=max(lens), max_seqlen_k=max(lens), is_causal=True)
    out0, lse0, _, _ = ops.fwd(q, k, v, **kw)
    out1, lse1, _, _ = ops.fwd(q, k, v, sm_margin=_margin(3), **kw)
    assert torch.equal(out0, out1) and torch.equal(lse0, lse1)
    for i, n in enumerate(lens):
        a = int(cu[i])
        check_fwd(out1[a:a + n].unsqueeze(0), None, q[a:a + n].unsqueeze(0), k[a:a + n].unsqueeze(0), v[a:a + n].unsqueeze(0),
                  causal=True, name=f"seq {i}")


@pytest.mark.parametrize("b,sq,sk,h,hk,d,causal,window", CASES[:4])
@pytest.mark.parametrize("ctas", [2, 6])
@pytest.mark.parametrize("deterministic", [False, True])
def test_backward_few_clusters(ops, b, sq, sk, h, hk, d, causal, window, ctas, deterministic):
    q, k, v = make_qkv(b, sq, sk, h, hk, d)
    dout = torch.randn_like(q)
    kw = dict(is_causal=causal, window_size_left=window[0], window_size_right=window[1])
    out, lse, _, _ = ops.fwd(q, k, v, **kw)
    def grads(sm_margin):   # FA3 convention: dq / dk / dv are written in place, the op returns softmax_d + scratch
        g = [torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)]
        ops.bwd(dout, q, k, v, out, lse, g[0], g[1], g[2], deterministic=deterministic, sm_margin=sm_margin, **kw)
        return g

    g0, g1 = grads(0), grads(_margin(ctas))
    assert torch.equal(g0[1], g1[1]) and torch.equal(g0[2], g1[2])   # dK, dV: accumulated by one CTA pair in TMEM
    if deterministic:
        assert torch.equal(g0[0], g1[0])
    tol = bwd_tolerances(q, k, v, dout, causal=causal, window=window)
    for name, g in zip(("dq", "dk", "dv"), g1):
        ref, bound = tol[name]
        err = (g.float().cpu() - ref).abs().max().item()
        assert err <= bound, f"{name}: {err:.3e} > {bound:.3e}"

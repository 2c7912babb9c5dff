"""Attention dropout (reference: dropout.h, philox.cuh, flash_api.cpp:505-517,990-1003).

The library's mask is a pure function of (rng_state, batch, head, query, key), restated on the CPU in
oracle/attention_oracle.py:philox_keep_mask.  These tests check, through the operator boundary,
  * S_dmask (return_softmax): sign pattern == the oracle's mask, magnitude == the oracle's probabilities,
  * forward output == oracle(dropout_mask) with the reference's 2x rule (tests/test_flash_attn.py:1121),
  * dq/dk/dv == fp32 autograd of the oracle with the same mask, 3x rule (:1130-1132),
  * the dropout fraction (:1125-1127) and that the generator offset advances by b*h*32 per call."""
import pytest
import torch

from oracle.attention_oracle import attention_oracle, philox_keep_mask  # checker only
from tests.util import bwd_tolerances, fwd_tolerances, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m
    from flash_attention_b200 import _lib

    _lib.load()
    return m


def _state(rng_state):
    seed, offset = (int(x) for x in rng_state.cpu().tolist())
    return seed & (2**64 - 1), offset


CASES = [
    # b, sq, sk, h, hk, d, causal, window, p, dtype
    (2, 128, 128, 4, 4, 64, False, (-1, -1), 0.17, torch.bfloat16),
    (2, 113, 203, 3, 3, 128, True, (-1, -1), 0.17, torch.float16),
    (1, 384, 500, 4, 2, 128, False, (100, 37), 0.3, torch.bfloat16),
    (1, 300, 300, 8, 2, 96, True, (-1, -1), 0.05, torch.bfloat16),
    (1, 257, 130, 2, 2, 40, True, (-1, -1), 0.5, torch.bfloat16),
    (1, 300, 421, 4, 2, 256, True, (-1, -1), 0.2, torch.bfloat16),     # head-dim 256 backward kernel
    (2, 200, 200, 2, 2, 192, False, (64, 64), 0.1, torch.float16),
]


@pytest.mark.parametrize("b,sq,sk,h,hk,d,causal,window,p,dtype", CASES)
def test_dropout_forward_backward(fab, b, sq, sk, h, hk, d, causal, window, p, dtype):
    from flash_attention_b200 import flash_attn_2_cuda as op

    q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype)
    do = torch.randn(b, sq, h, d, device="cuda", dtype=dtype)
    torch.cuda.manual_seed(1234)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    off0 = gen.get_offset()
    pad = (8 - d % 8) % 8
    qp, kp, vp, dop = (torch.nn.functional.pad(t, [0, pad]).contiguous() for t in (q, k, v, do))
    scale = d ** -0.5
    out, lse, s_dmask, rng_state = op.fwd(qp, kp, vp, None, None, p, scale, causal, window[0], window[1], 0.0, True, None)
    assert gen.get_offset() == off0 + b * h * 32
    seed, offset = _state(rng_state)
    assert offset == off0
    keep = philox_keep_mask(seed, offset, p, b, h, sq, sk)
    # ---- S_dmask: normalised probabilities, negative where dropped
    s = s_dmask[:, :, :sq, :sk].float().cpu()
    pref, _ = _probs(q, k, causal, window)
    nz = pref > 1e-4
    assert torch.equal((s >= 0)[nz], keep[nz]), "dropout mask differs from the oracle's Philox restatement"
    assert (s.abs() - pref).abs().max().item() < 2e-2 * pref.max().item() + 1e-3
    assert s_dmask[:, :, sq:, :].abs().max().item() == 0 if s_dmask.shape[2] > sq else True
    # ---- dropout fraction among visible entries
    vis = pref > 0
    frac = 1.0 - keep[vis].float().mean().item()
    assert abs(frac - p) < 0.02, frac
    # ---- forward
    kw = dict(causal=causal, window=window, dropout_p=p, dropout_mask=keep)
    ref32, lse_ref, bound = fwd_tolerances(q, k, v, **kw)
    err = (out[..., :d].float().cpu() - ref32).abs().max().item()
    assert err <= bound, f"out {err:.3e} > {bound:.3e}"
    fin = torch.isfinite(lse_ref)
    assert (lse.cpu()[fin] - lse_ref[fin]).abs().max().item() < 2e-3 * max(1.0, lse_ref[fin].abs().max().item())
    # ---- backward from the saved rng_state
    dq, dk, dv = torch.empty_like(qp), torch.empty_like(kp), torch.empty_like(vp)
    op.bwd(dop, qp, kp, vp, out, lse, dq, dk, dv, None, p, scale, causal, window[0], window[1], 0.0, False, None, rng_state)
    tol = bwd_tolerances(q, k, v, do, **kw)
    for name, g in (("dq", dq), ("dk", dk), ("dv", dv)):
        ref, bnd = tol[name]
        e = (g[..., :d].float().cpu() - ref).abs().max().item()
        assert not torch.isnan(g).any(), name
        assert e <= bnd, f"{name}: {e:.3e} > {bnd:.3e}"


def _probs(q, k, causal, window):
    """oracle probabilities (b, h, sq, sk) fp32 via V = identity columns: exp(s - lse) recomputed from the oracle's lse"""
    import math

    from oracle.attention_oracle import visible_mask

    qc, kc = q.float().cpu(), k.float().cpu()
    g = qc.shape[2] // kc.shape[2]
    kc = kc.repeat_interleave(g, dim=2)
    s = torch.einsum("bihd,bjhd->bhij", qc, kc) / math.sqrt(q.shape[-1])
    vis = visible_mask(q.shape[1], k.shape[1], causal, window, None, None, device="cpu")
    if vis is not None:
        s = s.masked_fill(~vis, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    return p, s


def test_dropout_autograd_and_reproducibility(fab):
    """public API: same generator state -> identical results; the autograd path carries rng_state to the backward"""
    b, s, h, d, p = 2, 256, 4, 128, 0.2
    q, k, v = make_qkv(b, s, s, h, h, d, torch.bfloat16)
    do = torch.randn_like(q)
    res = []
    for _ in range(2):
        torch.cuda.manual_seed(77)
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = fab.flash_attn_func(qq, kk, vv, dropout_p=p, causal=True)
        out.backward(do)
        res.append((out.detach(), qq.grad, kk.grad, vv.grad))
    for a, b_ in zip(*res):
        assert torch.equal(a, b_)
    out2 = fab.flash_attn_func(q, k, v, dropout_p=p, causal=True)  # generator advanced -> another mask
    assert not torch.equal(out2, res[0][0])
    # oracle check of the autograd path
    torch.cuda.manual_seed(77)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    keep = philox_keep_mask(gen.initial_seed(), gen.get_offset(), p, b, h, s, s)
    tol = bwd_tolerances(q, k, v, do, causal=True, dropout_p=p, dropout_mask=keep)
    for name, g in zip(("dq", "dk", "dv"), res[0][1:]):
        ref, bnd = tol[name]
        assert (g.float().cpu() - ref).abs().max().item() <= bnd, name


def test_dropout_varlen(fab):
    from flash_attention_b200 import flash_attn_2_cuda as op

    lens_q, lens_k = [70, 129, 5], [200, 129, 64]
    h, hk, d, p = 4, 2, 64, 0.25
    cu_q = torch.tensor([0] + list(torch.tensor(lens_q).cumsum(0)), dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0] + list(torch.tensor(lens_k).cumsum(0)), dtype=torch.int32, device="cuda")
    g = torch.Generator().manual_seed(3)
    q = torch.randn(sum(lens_q), h, d, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(sum(lens_k), hk, d, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(sum(lens_k), hk, d, generator=g).to(torch.bfloat16).cuda()
    do = torch.randn_like(q)
    torch.cuda.manual_seed(5)
    out, lse, s_dmask, rng_state = op.varlen_fwd(q, k, v, None, cu_q, cu_k, None, None, None, None, max(lens_q), max(lens_k),
                                                 p, d ** -0.5, False, True, -1, -1, 0.0, True, None)
    seed, offset = _state(rng_state)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    op.varlen_bwd(do, q, k, v, out, lse, dq, dk, dv, cu_q, cu_k, None, max(lens_q), max(lens_k), p, d ** -0.5, False, True,
                  -1, -1, 0.0, False, None, rng_state)
    keep_all = philox_keep_mask(seed, offset, p, len(lens_q), h, max(lens_q), max(lens_k))
    for i, (lq, lk) in enumerate(zip(lens_q, lens_k)):
        qs, ks = slice(int(cu_q[i]), int(cu_q[i + 1])), slice(int(cu_k[i]), int(cu_k[i + 1]))
        keep = keep_all[i : i + 1, :, :lq, :lk]
        kw = dict(causal=True, dropout_p=p, dropout_mask=keep)
        qi, ki, vi, doi = q[qs][None], k[ks][None], v[ks][None], do[qs][None]
        ref32, _, bound = fwd_tolerances(qi, ki, vi, **kw)
        assert (out[qs][None].float().cpu() - ref32).abs().max().item() <= bound, f"seq {i} out"
        sd = s_dmask[i, :, :lq, :lk].float().cpu()
        pref, _ = _probs(qi, ki, True, (-1, -1))
        nz = pref[0] > 1e-4
        assert torch.equal((sd >= 0)[nz], keep[0][nz]), f"seq {i} mask"
        tol = bwd_tolerances(qi, ki, vi, doi, **kw)
        for name, gq in (("dq", dq[qs]), ("dk", dk[ks]), ("dv", dv[ks])):
            ref, bnd = tol[name]
            assert (gq[None].float().cpu() - ref).abs().max().item() <= bnd, f"seq {i} {name}"


def test_dropout_hd256_forward(fab):
    from flash_attention_b200 import flash_attn_2_cuda as op

    b, s, h, d, p = 1, 300, 2, 256, 0.1
    q, k, v = make_qkv(b, s, s, h, h, d, torch.bfloat16)
    torch.cuda.manual_seed(11)
    out, lse, _, rng_state = op.fwd(q, k, v, None, None, p, d ** -0.5, False, -1, -1, 0.0, False, None)
    seed, offset = _state(rng_state)
    keep = philox_keep_mask(seed, offset, p, b, h, s, s)
    ref32, _, bound = fwd_tolerances(q, k, v, dropout_p=p, dropout_mask=keep)
    assert (out.float().cpu() - ref32).abs().max().item() <= bound


def test_dropout_argument_errors(fab):
    from flash_attention_b200 import flash_attn_2_cuda as op

    q, k, v = make_qkv(1, 128, 128, 2, 2, 64, torch.bfloat16)
    with pytest.raises(RuntimeError, match="Softcapping does not support dropout"):
        op.fwd(q, k, v, None, None, 0.1, 0.125, False, -1, -1, 30.0, False, None)
    with pytest.raises(RuntimeError, match="return_softmax is only supported when p_dropout > 0.0"):
        op.fwd(q, k, v, None, None, 0.0, 0.125, False, -1, -1, 0.0, True, None)
    with pytest.raises(RuntimeError, match="p_dropout"):
        op.fwd(q, k, v, None, None, 1.0, 0.125, False, -1, -1, 0.0, False, None)


@pytest.mark.parametrize("b,sq,sk,h,hk,d,causal,p", [
    (2, 256, 256, 4, 4, 128, False, 0.17),
    (1, 384, 500, 4, 2, 128, True, 0.3),      # GQA, sq != sk, causal (bottom-right aligned)
    (2, 130, 257, 3, 3, 64, False, 0.17),     # ragged sizes, head dim 64
    (1, 200, 333, 2, 1, 96, True, 0.5),
])
def test_dropout_mask_is_the_reference_kernels_mask(fab, b, sq, sk, h, hk, d, causal, p):
    """Same CUDA generator state -> the SAME elements dropped as by the reference's FA2 kernels (the flash_attn wheel in
    the image): the dropout stream is the reference's (dropout.h:20-47), not a look-alike.  FA2 returns S_dmask with a
    negative sign where an element was dropped ("for testing only", flash_attn_interface.py:1196), in plain (query, key)
    layout inside the 128-rounded array (tests/test_flash_attn.py:411-463), so sign patterns are directly comparable;
    and with identical masks the outputs and gradients of the two kernels agree like two bf16 kernels do."""
    from flash_attention_b200 import flash_attn_2_cuda as op
    from tests.util import load_reference_fa2

    real = load_reference_fa2()
    q, k, v = make_qkv(b, sq, sk, h, hk, d, torch.bfloat16)
    do = torch.randn(b, sq, h, d, device="cuda", dtype=torch.bfloat16)
    scale = d ** -0.5
    res = {}
    for name, mod in (("ref", real), ("ours", op)):
        torch.cuda.manual_seed(4321)   # same (seed, offset) for both
        out, lse, s_dmask, rng = mod.fwd(q, k, v, None, None, p, scale, causal, -1, -1, 0.0, True, None)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        mod.bwd(do, q, k, v, out, lse, dq, dk, dv, None, p, scale, causal, -1, -1, 0.0, False, None, rng)
        res[name] = (out, s_dmask[:, :, :sq, :sk].float(), rng.clone(), dq, dk, dv)
    assert torch.equal(res["ref"][2], res["ours"][2])   # rng_state (seed, offset) as returned
    pref, _ = _probs(q, k, causal, (-1, -1))
    nz = (pref > 1e-4).cuda()   # entries whose probability survives fp16/bf16 storage in both kernels (sign is meaningful)
    ref_keep, our_keep = res["ref"][1] >= 0, res["ours"][1] >= 0
    assert nz.float().mean().item() > 0.05
    assert torch.equal(ref_keep[nz], our_keep[nz]), "dropped elements differ from the reference kernels'"
    frac = 1.0 - our_keep[nz].float().mean().item()
    assert abs(frac - p) < 0.03, frac
    for name, i, tol in (("out", 0, 2.0 ** -6), ("dq", 3, 2.0 ** -5), ("dk", 4, 2.0 ** -5), ("dv", 5, 2.0 ** -5)):
        a, r = res["ours"][i].float(), res["ref"][i].float()
        assert (a - r).abs().max().item() <= tol * max(1.0, r.abs().max().item()), name

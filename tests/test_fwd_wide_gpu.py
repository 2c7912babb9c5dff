"""The wide-S forward kernel (fa_fwd_wide_sm100.cuh: S = Q K^T 256 keys at a time, two softmax warpgroups per query tile,
QK(g+1) issued before PV(g)), forced with FA_B200_FWD_WIDE=1, against the oracle with the reference's tolerance rule
(tests/test_flash_attn.py:1121): masks, ragged lengths, sq != sk, GQA, head dims below 128, fp16, reference jumps in late
blocks (lazy rescale across the two groups), more work items than CTAs (persistent loop, dynamic draws), and -- bit for bit
-- against the two-tile kernel's LSE / output tolerance at a BASELINE shape slice."""
import os

import pytest
import torch

from tests.util import check_fwd, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m

    return m


@pytest.fixture(autouse=True)
def _wide(monkeypatch):
    monkeypatch.setenv("FA_B200_FWD_WIDE", "1")


def _run(fab, q, k, v, **kw):
    out, lse, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True, **kw)
    check_fwd(out, lse, q, k, v, **{("window" if a == "window_size" else a): b for a, b in kw.items()})
    return out, lse


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("sq,sk", [(256, 256), (512, 1024), (1024, 512), (384, 640), (300, 777), (1000, 1000), (129, 2049)])
def test_shapes(fab, causal, sq, sk):
    q, k, v = make_qkv(2, sq, sk, 4, 4, 128)
    _run(fab, q, k, v, causal=causal)


@pytest.mark.parametrize("window", [(128, 0), (200, 50), (0, 300), (1000, 0)])
def test_local_windows(fab, window):
    q, k, v = make_qkv(1, 900, 1100, 4, 4, 128)
    _run(fab, q, k, v, window_size=window)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("d", [128, 96, 72])
def test_dtypes_and_head_dims(fab, dtype, d):
    q, k, v = make_qkv(1, 640, 896, 4, 2, d, dtype)   # GQA 2:1
    _run(fab, q, k, v, causal=True)


def test_gqa_8_to_1(fab):
    q, k, v = make_qkv(1, 512, 768, 16, 2, 128)
    _run(fab, q, k, v)


@pytest.mark.parametrize("cols", [(0, 64), (100, 164), (300, 301), (640, 704), (1500, 1536)])
@pytest.mark.parametrize("gain", [6.0, 40.0])
def test_late_blocks_outgrow_the_reference(fab, cols, gain):
    # keys are streamed right to left: early keys come last and force a rescale of O in both column halves
    q, k, v = make_qkv(1, 512, 1536, 2, 2, 128)
    k = k.clone()
    k[:, cols[0]:cols[1]] *= gain
    _run(fab, q, k, v)


def test_fully_masked_rows_and_tiles(fab):
    # causal with sq > sk: the first sq - sk query rows see no key at all (zeros, LSE = +inf), whole tiles included
    q, k, v = make_qkv(1, 768, 300, 2, 2, 128)
    _run(fab, q, k, v, causal=True)


@pytest.mark.parametrize("dyn", ["0", "2"])
def test_more_items_than_ctas(fab, dyn, monkeypatch):
    monkeypatch.setenv("FA_B200_SCHED_DYNAMIC", dyn)
    q, k, v = make_qkv(3, 1536, 1536, 16, 4, 128)   # 3 * 16 * 12 = 576 items on 148 CTAs, uneven causal lengths
    _run(fab, q, k, v, causal=True)


def test_matches_two_tile_kernel_at_baseline_slice(fab, monkeypatch):
    q, k, v = make_qkv(1, 8192, 8192, 2, 2, 128)
    out_w, lse_w, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True)
    monkeypatch.setenv("FA_B200_FWD_WIDE", "0")
    out_t, lse_t, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True)
    assert (out_w.float() - out_t.float()).abs().max().item() <= 2e-3
    assert (lse_w - lse_t).abs().max().item() <= 1e-4


# ---- one-pass block pairs (exponentials against the row maximum of the blocks BEFORE the pair; fa_fwd_wide_sm100.cuh)
@pytest.mark.parametrize("kval,col", [(300.0, 13), (300.0, 700), (450.0, 200), (700.0, 5), (700.0, 1000), (1500.0, 333)])
def test_onepass_huge_jumps(fab, kval, col):
    # every query shares a component with ONE key, whose score is 4 * kval / sqrt(128) nats above the rest: 153 / 230 / 357 /
    # 765 in the exp2 domain -- below, inside and beyond the range the lagging reference can absorb (2^214): the last ones
    # overflow an accumulator and go through the exact per-row fallback
    b, s, h, d = 1, 1536, 2, 128
    q, k, v = make_qkv(b, s, s, h, h, d)
    q = q.clone(); k = k.clone()
    q[..., 0] = 4.0
    k[:, col, :, 0] = kval
    if kval > 400.0:
        _run(fab, q, k, v)      # exact fallback: the reference's rule holds even for these one-hot rows (the bound is ~1e-7)
        return
    # Below the fallback's range a one-pass block pair keeps the reference it came with, so the dominant probability is
    # bf16(2^frac) instead of exactly 1 (as after any lazy rescale, here for jumps of any size): at most one bf16 ulp of O on
    # (near-)one-hot rows.  The reference's rule degenerates to atol = 1e-7 on such rows (its low-precision oracle is exact).
    from tests.util import fwd_tolerances

    out, lse, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True)
    ref32, lse_ref, bound = fwd_tolerances(q, k, v)
    err = (out.float().cpu() - ref32).abs()
    assert not torch.isnan(out).any()
    assert (err <= torch.clamp(ref32.abs() * 2.0 ** -7, min=max(bound, 1e-6))).all(), f"max err {err.max().item():.3e}"
    assert (lse.float().cpu() - lse_ref).abs().max().item() <= 2e-3 * lse_ref.abs().max().item()


@pytest.mark.parametrize("causal", [False, True])
def test_onepass_strongly_negative_scores(fab, causal):
    # every score is ~ -250 nats: a reference of 0 would underflow every exponential
    b, s, h, d = 1, 1280, 2, 128
    q, k, v = make_qkv(b, s, s, h, h, d)
    q = q.clone(); k = k.clone()
    q[..., 1] = 50.0
    k[..., 1] = -56.0
    _run(fab, q, k, v, causal=causal)


def test_onepass_steady_drift(fab):
    # every earlier key block scores higher than the later ones: the lagging reference moves at every block pair
    b, s, h, d = 2, 4096, 2, 128
    q, k, v = make_qkv(b, s, s, h, h, d)
    ramp = torch.linspace(6.0, 0.2, s, device="cuda").view(1, s, 1, 1)
    k = (k.float() * ramp).to(torch.bfloat16)
    _run(fab, q, k, v, causal=True)

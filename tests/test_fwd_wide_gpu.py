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

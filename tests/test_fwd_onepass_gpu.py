"""The forward's one-pass blocks (exponentials against the stale reference maximum, fa_fwd_sm100.cuh `opt` / fix_half):
inputs built so that the row maxima jump by many orders of magnitude in blocks that are processed LATE (the kernel streams
the key blocks right to left, so early keys come last), in the first half / second half / polynomial columns of a block
only, and beyond 2^127 relative to the running reference.  All against the oracle with the reference's tolerance rule."""
import pytest
import torch

from tests.util import check_fwd, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m

    return m


def _run(fab, q, k, v, **kw):
    out, lse, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True, **kw)
    check_fwd(out, lse, q, k, v, **{("window" if a == "window_size" else a): b for a, b in kw.items()})


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cols", [(0, 128), (0, 64), (64, 128), (12, 16), (76, 80), (300, 301), (1021, 1024)])
@pytest.mark.parametrize("gain", [6.0, 40.0])
def test_late_blocks_outgrow_the_reference(fab, dtype, cols, gain):
    b, s, h, d = 1, 1024, 2, 128
    q, k, v = make_qkv(b, s, s, h, h, d, dtype)
    k = k.clone()
    k[:, cols[0]:cols[1]] *= gain       # those keys' scores are `gain` times larger: exp() of them dwarfs everything seen so far
    _run(fab, q, k, v)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("col", [5, 13, 70, 127, 140, 525])
def test_jump_beyond_2_pow_127(fab, dtype, col):
    b, s, h, d = 1, 768, 2, 128
    q, k, v = make_qkv(b, s, s, h, h, d, dtype)
    q = q.clone(); k = k.clone()
    # every query shares a component with ONE key: its score is 4 * 300 / sqrt(128) = 106 nats above the rest, i.e. more
    # than 2^127 in the exp2 domain -- exp2 against the stale reference overflows to inf (MUFU columns) or wraps around
    # (polynomial columns, col % 16 >= 12)
    q[..., 0] = 4.0
    k[:, col, :, 0] = 300.0
    _run(fab, q, k, v)


@pytest.mark.parametrize("causal", [False, True])
def test_increasing_scores_toward_early_keys(fab, causal):
    # a steady drift: every earlier key block scores higher than the later ones, so the guard fires repeatedly
    b, s, h, d = 2, 2048, 2, 128
    q, k, v = make_qkv(b, s, s, h, h, d)
    k = k.clone()
    ramp = torch.linspace(6.0, 0.2, s, device="cuda").view(1, s, 1, 1)
    k = (k.float() * ramp).to(torch.bfloat16)
    _run(fab, q, k, v, causal=causal)


def test_gqa_packed_rows_and_two_tiles(fab):
    b, sq, sk, h, hk, d = 1, 640, 1536, 8, 2, 128
    q, k, v = make_qkv(b, sq, sk, h, hk, d)
    k = k.clone()
    k[:, 100:164] *= 25.0
    _run(fab, q, k, v, causal=True)


@pytest.mark.parametrize("d", [64, 256])
def test_other_head_dims(fab, d):
    b, s, h = 1, 900, 2
    q, k, v = make_qkv(b, s, s, h, h, d)
    k = k.clone()
    k[:, 10:40] *= 20.0
    _run(fab, q, k, v)

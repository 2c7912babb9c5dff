"""FP8 e4m3 forward (BASELINE C4 family).  Oracle: fp32 attention on the de-quantised inputs; tolerance in the
spirit of the reference's fp8 tests (tests/cute/test_flash_attn.py:2119-2123, 2389-2406: mult=4 on the
quantised-P reference error, cosine > 0.99): |out - ref| <= 4 * |ref_with_e4m3_P - ref| + 2e-2."""
import pytest
import torch

from oracle.attention_oracle import attention_oracle  # checker only
from tests.util import make_qkv

pytestmark = pytest.mark.gpu
F8 = torch.float8_e4m3fn


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m
    from flash_attention_b200 import _lib

    _lib.load()
    return m


@pytest.mark.parametrize("b,sq,sk,h,hk,causal,window,descale", [
    (1, 128, 128, 1, 1, False, (-1, -1), False),
    (2, 384, 640, 4, 2, False, (-1, -1), True),
    (1, 512, 512, 2, 2, True, (-1, -1), True),
    (1, 1024, 1024, 2, 2, True, (256, 0), False),   # sliding window, the C4 mask
    (2, 200, 333, 3, 3, False, (-1, -1), True),
])
def test_fp8_forward(fab, b, sq, sk, h, hk, causal, window, descale):
    q, k, v = make_qkv(b, sq, sk, h, hk, 128)
    q8, k8, v8 = q.to(F8), k.to(F8), v.to(F8)
    ds = [(torch.rand(b, hk, device="cuda") + 0.5).float() if descale else None for _ in range(3)]
    out, lse, _ = fab.flash_attn_func(q8, k8, v8, causal=causal, window_size=window, return_attn_probs=True,
                                     q_descale=ds[0], k_descale=ds[1], v_descale=ds[2])
    assert out.dtype == torch.bfloat16
    c = lambda t: None if t is None else t.cpu()
    kw = dict(causal=causal, window=window, q_descale=c(ds[0]), k_descale=c(ds[1]), v_descale=c(ds[2]))
    qf, kf, vf = q8.cpu().float(), k8.cpu().float(), v8.cpu().float()
    ref, lse_ref = attention_oracle(qf, kf, vf, **kw)
    refq, _ = attention_oracle(qf, kf, vf, p_dtype=F8, **kw)
    err = (out.float().cpu() - ref).abs().max().item()
    assert not torch.isnan(out).any()
    assert err <= 4 * (refq - ref).abs().max().item() + 2e-2, err
    cos = torch.nn.functional.cosine_similarity(out.float().cpu().flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.99
    fin = torch.isfinite(lse_ref)
    assert (lse.cpu()[fin] - lse_ref[fin]).abs().max().item() < 5e-3


def test_fp8_preserves_tail_mass(fab):
    """P is up-scaled by 2^8 before the e4m3 cast, so keys whose probability is ~2^-12 still contribute
    (reference KAT: test_flash_attn_fp8_paged_decode_preserves_tail_mass, tests/cute/test_flash_attn.py:2389-2406)."""
    b, sq, sk, h, d = 1, 128, 4096, 1, 128
    q = torch.zeros(b, sq, h, d, device="cuda")
    k = torch.zeros(b, sk, h, d, device="cuda")
    v = torch.zeros(b, sk, h, d, device="cuda")
    q[..., 0] = 8.0
    k[:, 0, :, 0] = 12.0                             # one dominant key: logit 8*12/sqrt(128) = 8.49 above the rest
    # -> tail probabilities are e^-8.49 = 2.1e-4 of the max: below e4m3's smallest subnormal (2^-9) unless P is up-scaled
    v[:, 0] = 0.0
    v[:, 1:] = 1.0                                   # all the mass that matters sits in the 4095-key tail
    out = fab.flash_attn_func(q.to(F8), k.to(F8), v.to(F8))
    ref, _ = attention_oracle(q.to(F8).cpu().float(), k.to(F8).cpu().float(), v.to(F8).cpu().float())
    assert torch.allclose(out.float().cpu(), ref, atol=0.01, rtol=0.1)

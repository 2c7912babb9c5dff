"""GPU parity tests of the forward path (through the public API -> operator module -> C ABI -> sm_100a
kernels) against the CPU oracle, with the reference's tolerance rule (tests/util.py)."""
import math

import pytest
import torch

from tests.util import check_fwd, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m
    from flash_attention_b200 import _lib

    _lib.load()
    assert torch.cuda.is_available()
    return m


CASES = [
    # b, sq, sk, h, hk, d, causal, window, softcap, dtype
    (2, 128, 128, 4, 4, 64, False, (-1, -1), 0.0, torch.bfloat16),   # BASELINE C1 shape (d=64)
    (2, 128, 128, 4, 4, 64, True, (-1, -1), 0.0, torch.float16),
    (1, 256, 256, 2, 2, 128, False, (-1, -1), 0.0, torch.bfloat16),
    (2, 113, 203, 3, 3, 128, False, (-1, -1), 0.0, torch.bfloat16),  # reference seqlen pairs
    (2, 113, 203, 3, 3, 128, True, (-1, -1), 0.0, torch.bfloat16),
    (1, 1023, 1024, 2, 1, 128, True, (-1, -1), 0.0, torch.bfloat16),
    (1, 512, 256, 2, 2, 128, True, (-1, -1), 0.0, torch.float16),    # sq > sk: dead rows
    (1, 384, 500, 4, 2, 128, False, (100, 37), 0.0, torch.bfloat16),
    (1, 384, 500, 2, 2, 96, False, (64, -1), 0.0, torch.bfloat16),
    (1, 700, 700, 1, 1, 128, True, (129, -1), 0.0, torch.bfloat16),
    (1, 256, 256, 2, 2, 128, False, (-1, -1), 15.0, torch.bfloat16),
    (2, 1, 777, 8, 1, 128, False, (-1, -1), 0.0, torch.bfloat16),    # decode, packed GQA
    (2, 5, 260, 16, 2, 128, True, (-1, -1), 0.0, torch.bfloat16),
    (1, 130, 257, 2, 2, 32, True, (-1, -1), 0.0, torch.bfloat16),
    (1, 130, 257, 2, 2, 40, False, (-1, -1), 0.0, torch.float16),
]


@pytest.mark.parametrize("b,sq,sk,h,hk,d,causal,window,softcap,dtype", CASES)
def test_flash_attn_func_matches_oracle(fab, b, sq, sk, h, hk, d, causal, window, softcap, dtype):
    q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype)
    out, lse, _ = fab.flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, return_attn_probs=True)
    check_fwd(out, lse, q, k, v, causal=causal, window=window, softcap=softcap)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("per_batch", [False, True])
def test_alibi(fab, causal, per_batch):
    b, sq, sk, h = 2, 200, 333, 4
    q, k, v = make_qkv(b, sq, sk, h, h, 128)
    sl = (torch.rand((b, h) if per_batch else (h,), device="cuda") * 0.3).float()
    out, lse, _ = fab.flash_attn_func(q, k, v, causal=causal, alibi_slopes=sl, return_attn_probs=True)
    check_fwd(out, lse, q, k, v, causal=causal, alibi_slopes=sl)


def test_strided_inputs_and_out_layout(fab):
    """qkv-packed views: non-contiguous row/head strides go straight into the TMA descriptors."""
    b, s, h, d = 2, 300, 4, 128
    qkv = torch.randn(b, s, 3, h, d, device="cuda", dtype=torch.bfloat16)
    out = fab.flash_attn_qkvpacked_func(qkv, causal=True)
    check_fwd(out, None, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=True)


def test_empty_keys_gives_zero_out_inf_lse(fab):
    q, k, v = make_qkv(1, 64, 0, 2, 2, 128)
    out, lse, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True)
    assert torch.all(out == 0) and torch.all(torch.isinf(lse) & (lse > 0))


def test_error_messages(fab):
    q, k, v = make_qkv(1, 64, 64, 3, 2, 128)
    with pytest.raises(RuntimeError, match="must divide"):
        fab.flash_attn_func(q, k, v)
    q, k, v = make_qkv(1, 64, 64, 2, 2, 128, torch.float32)
    with pytest.raises(RuntimeError, match="fp16 and bf16"):
        fab.flash_attn_func(q, k, v)


def test_key_permutation_invariance_full_size(fab):
    """Size-independent property at a BASELINE-scale shape (non-causal): permuting keys/values together
    leaves the output unchanged up to accumulation order; compared on the device."""
    q, k, v = make_qkv(1, 2048, 8192, 4, 4, 128)
    perm = torch.randperm(8192, device="cuda")
    o1 = fab.flash_attn_func(q, k, v)
    o2 = fab.flash_attn_func(q, k[:, perm], v[:, perm])
    assert (o1.float() - o2.float()).abs().max().item() < 4e-3


def test_full_size_rows_vs_oracle(fab):
    """BASELINE C2 shape (b4 h16 s8192 d128), oracle evaluated on a slice: last 256 query rows of one
    (batch, head) against all 8192 keys (bottom-right aligned => valid for causal too)."""
    b, h, s, d = 4, 16, 8192, 128
    q, k, v = make_qkv(b, s, s, h, h, d, seed=1)
    for causal in (False, True):
        out, lse, _ = fab.flash_attn_func(q, k, v, causal=causal, return_attn_probs=True)
        bi, hi = 3, 11
        check_fwd(out[bi:bi + 1, -256:, hi:hi + 1], lse[bi:bi + 1, hi:hi + 1, -256:], q[bi:bi + 1, -256:, hi:hi + 1],
                  k[bi:bi + 1, :, hi:hi + 1], v[bi:bi + 1, :, hi:hi + 1], causal=causal)


def test_dead_and_live_rows_in_one_tile(fab):
    """sq > sk, causal: rows 0..sq-sk-1 see no key.  When dead and live rows share a 128-row tile the dead rows
    must still come out as exact zeros with lse = +inf (regression: the polynomial exp2 returned 2^-125 instead of 0
    for -inf scores, which gave fully masked rows a non-zero row sum)."""
    q, k, v = make_qkv(1, 256, 100, 2, 2, 128)
    out, lse, _ = fab.flash_attn_func(q, k, v, causal=True, return_attn_probs=True)
    assert torch.all(out[:, :156] == 0) and torch.all(torch.isinf(lse[:, :, :156]))
    check_fwd(out, lse, q, k, v, causal=True)


@pytest.mark.parametrize("d", [160, 192, 224, 256])
@pytest.mark.parametrize("causal", [False, True])
def test_forward_head_dims_up_to_256(fab, d, causal):
    """head dims 129..256 run the 1-tile forward configuration (S 128 + O 256 TMEM columns)."""
    q, k, v = make_qkv(2, 300, 421, 4, 2, d)
    out, lse, _ = fab.flash_attn_func(q, k, v, causal=causal, return_attn_probs=True)
    check_fwd(out, lse, q, k, v, causal=causal)


def test_forward_is_bitwise_reproducible(fab):
    """Reference: test_flash_attn_race_condition (tests/test_flash_attn.py:2199): repeated runs must be bit-equal."""
    q, k, v = make_qkv(8, 500, 777, 6, 2, 128)
    o0, l0, _ = fab.flash_attn_func(q, k, v, causal=True, return_attn_probs=True)
    for _ in range(50):
        o, l, _ = fab.flash_attn_func(q, k, v, causal=True, return_attn_probs=True)
        assert torch.equal(o, o0) and torch.equal(l, l0)


def test_varlen_with_empty_sequences(fab):
    """cu_seqlens with zero-length query and key sequences (reference varlen tests draw random lengths incl. 0)."""
    from oracle.attention_oracle import attention_oracle_varlen

    h, hk, d = 4, 2, 128
    lens_q, lens_k = [0, 130, 64, 5], [77, 0, 200, 5]
    cu_q = torch.tensor([0, 0, 130, 194, 199], dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0, 77, 77, 277, 282], dtype=torch.int32, device="cuda")
    q = torch.randn(199, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(282, hk, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(282, hk, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    out, lse, _ = fab.flash_attn_varlen_func(q, k, v, cu_q, cu_k, 130, 200, causal=False, return_attn_probs=True)
    ref, lse_ref = attention_oracle_varlen(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), cu_q.cpu(), cu_k.cpu(), 130, 200)
    assert (out.float().cpu() - ref.float()).abs().max().item() <= 2e-2
    assert torch.all(out[0:130] == 0) and torch.all(torch.isinf(lse[:, 0:130]))  # sequence 1 has no keys
    out.backward(torch.randn_like(out))
    assert torch.isfinite(q.grad).all() and torch.isfinite(k.grad).all() and torch.isfinite(v.grad).all()
    assert torch.all(q.grad[0:130] == 0) and torch.all(k.grad[0:77] == 0)  # no partner -> zero gradient


@pytest.mark.parametrize("causal", [False, True])
def test_varlen_forward_leftpad_k(causal):
    """mha_varlen_fwd's leftpad_k (flash_api.cpp:740-747, block_info.h:12-45): keys of batch b start leftpad_k[b] rows
    into its packed rows and seqused_k counts from the start of those rows, i.e. batch b attends to rows
    [cu_k[b] + leftpad[b], cu_k[b] + seqused_k[b])."""
    from flash_attention_b200 import flash_attn_2_cuda as op
    from oracle.attention_oracle import attention_oracle

    h, hk, d = 4, 2, 128
    lens_q, cap_k = [70, 128, 33], [300, 260, 129]          # packed rows reserved per batch
    leftpad, used = [17, 0, 128], [250, 260, 129]           # seqused_k includes the left pad
    cu_q = torch.tensor([0, 70, 198, 231], dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0, 300, 560, 689], dtype=torch.int32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(231, h, d, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(689, hk, d, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(689, hk, d, device="cuda", dtype=torch.bfloat16, generator=g)
    lp = torch.tensor(leftpad, dtype=torch.int32, device="cuda")
    su = torch.tensor(used, dtype=torch.int32, device="cuda")
    out, lse, _, _ = op.varlen_fwd(q, k, v, None, cu_q, cu_k, su, lp, None, None, max(lens_q), max(cap_k), 0.0, d ** -0.5,
                                   False, causal, -1, -1, 0.0, False, None)
    for b in range(3):
        qs = q[int(cu_q[b]):int(cu_q[b + 1])].unsqueeze(0).cpu()
        lo, hi = int(cu_k[b]) + leftpad[b], int(cu_k[b]) + used[b]
        ref, lse_ref = attention_oracle(qs, k[lo:hi].unsqueeze(0).cpu(), v[lo:hi].unsqueeze(0).cpu(), causal=causal)
        got = out[int(cu_q[b]):int(cu_q[b + 1])].float().cpu()
        assert (got - ref[0].float()).abs().max().item() <= 2e-2, b
        got_lse = lse[:, int(cu_q[b]):int(cu_q[b + 1])].cpu()
        fin = torch.isfinite(lse_ref[0])
        assert torch.equal(torch.isfinite(got_lse), fin)
        assert (got_lse[fin] - lse_ref[0][fin]).abs().max().item() <= 2e-3 * 8

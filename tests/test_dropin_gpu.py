"""Drop-in test: the REFERENCE's own flash_attn/flash_attn_interface.py (the copy installed in the image,
flash_attn 2.8.x) runs unchanged on top of this package's operator module, and its results match the
reference's own FA2 kernels (the installed flash_attn_2_cuda extension) on identical inputs.

Tolerance: both are 16-bit kernels with fp32 accumulation; |ours - FA2| <= 2 * |FA2 - fp32 oracle| + 1e-3 is
not needed -- a plain absolute bound of 2 bf16 ulps of the output magnitude (out ~ O(1): 1.6e-2) is asserted,
plus the oracle rule through tests/util.py."""
import importlib
import importlib.util
import os
import sys

import pytest
import torch

from tests.util import check_fwd, load_reference_fa2, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack():
    spec = importlib.util.find_spec("flash_attn")
    if spec is None:
        pytest.skip("reference flash_attn package is not installed in this image")
    iface_path = os.path.join(list(spec.submodule_search_locations)[0], "flash_attn_interface.py")
    real = load_reference_fa2()  # the reference's compiled kernels
    import flash_attention_b200 as fab

    sys.modules.pop("flash_attn_2_cuda", None)
    ours = fab.install()  # sys.modules['flash_attn_2_cuda'] = our operator module
    s = importlib.util.spec_from_file_location("ref_iface_on_fa_b200", iface_path)
    iface = importlib.util.module_from_spec(s)
    s.loader.exec_module(iface)  # the reference file, unmodified
    assert iface.flash_attn_gpu is ours
    return iface, real


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("hk", [4, 1])
def test_reference_interface_forward_backward(stack, causal, hk):
    iface, real = stack
    b, s, h, d = 2, 384, 4, 128
    q, k, v = make_qkv(b, s, s, h, hk, d)
    do = torch.randn_like(q)
    q1, k1, v1 = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = iface.flash_attn_func(q1, k1, v1, causal=causal)  # reference autograd.Function + custom_op on our kernels
    out.backward(do)
    check_fwd(out.detach(), None, q, k, v, causal=causal)
    # the reference's own kernels on the same tensors
    scale = d ** -0.5
    r_out, r_lse, _, _ = real.fwd(q, k, v, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
    assert (out.detach().float() - r_out.float()).abs().max().item() <= 1.6e-2
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    real.bwd(do, q, k, v, r_out, r_lse, dq, dk, dv, None, 0.0, scale, causal, -1, -1, 0.0, False, None, None)
    for g, r in ((q1.grad, dq), (k1.grad, dk), (v1.grad, dv)):
        assert (g.float() - r.float()).abs().max().item() <= 3e-2 * max(1.0, r.float().abs().max().item())


def test_reference_interface_varlen_and_kvcache(stack):
    iface, real = stack
    h, hk, d = 4, 2, 128
    lens = [100, 256, 31]
    cu = torch.tensor([0, 100, 356, 387], dtype=torch.int32, device="cuda")
    q = torch.randn(387, h, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(387, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(387, hk, d, device="cuda", dtype=torch.bfloat16)
    out = iface.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True)
    scale = d ** -0.5
    r = real.varlen_fwd(q, k, v, None, cu, cu, None, None, None, None, max(lens), max(lens), 0.0, scale, False, True, -1, -1,
                        0.0, False, None)
    assert (out.float() - r[0].float()).abs().max().item() <= 1.6e-2
    # decode against a kv cache with per-batch lengths
    b, sk = 3, 1024
    qd = torch.randn(b, 1, h, d, device="cuda", dtype=torch.bfloat16)
    kc = torch.randn(b, sk, hk, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(b, sk, hk, d, device="cuda", dtype=torch.bfloat16)
    cs = torch.tensor([1000, 17, 512], dtype=torch.int32, device="cuda")
    o = iface.flash_attn_with_kvcache(qd, kc, vc, cache_seqlens=cs, causal=True)
    r = real.fwd_kvcache(qd, kc, vc, None, None, cs, None, None, None, None, None, None, None, scale, True, -1, -1, 0.0, True, 0)
    assert (o.float() - r[0].float()).abs().max().item() <= 1.6e-2

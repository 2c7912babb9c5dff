"""FA3-style operator surface mapped on the same kernels: pack_gqa / num_splits knobs, seqused_q/k,
page_table, public combine, scheduler metadata."""
import pytest
import torch

from oracle.attention_oracle import attention_oracle  # checker only
from tests.util import check_fwd, make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fa3():
    import flash_attention_b200 as m

    return m.flash_attn_3


@pytest.mark.parametrize("pack_gqa", [None, True, False])
@pytest.mark.parametrize("num_splits", [1, 3])
def test_knobs_do_not_change_results(fa3, pack_gqa, num_splits):
    q, k, v = make_qkv(2, 40, 700, 8, 2, 128)
    out, lse = fa3.flash_attn_func(q, k, v, causal=True, pack_gqa=pack_gqa, num_splits=num_splits, return_attn_probs=True)
    check_fwd(out, lse, q, k, v, causal=True)


def test_seqused_q_and_k(fa3):
    torch.manual_seed(0)
    h, hk, d = 4, 2, 128
    cu = torch.tensor([0, 300, 556], dtype=torch.int32, device="cuda")
    q = torch.randn(556, h, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(556, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(556, hk, d, device="cuda", dtype=torch.bfloat16)
    uq = torch.tensor([200, 256], dtype=torch.int32, device="cuda")
    uk = torch.tensor([300, 100], dtype=torch.int32, device="cuda")
    out = fa3.flash_attn_varlen_func(q, k, v, cu, cu, 300, 300, seqused_q=uq, seqused_k=uk, causal=True)
    for i in range(2):
        a = int(cu[i])
        nq, nk = int(uq[i]), int(uk[i])
        check_fwd(out[a:a + nq].unsqueeze(0), None, q[a:a + nq].unsqueeze(0), k[a:a + nk].unsqueeze(0), v[a:a + nk].unsqueeze(0),
                  causal=True)


def test_scheduler_metadata_and_kvcache(fa3):
    b, h, hk, d, page = 4, 16, 2, 128, 256
    nblk = 8
    kc = torch.randn(b * nblk, page, hk, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(b * nblk, page, hk, d, device="cuda", dtype=torch.bfloat16)
    pt = torch.randperm(b * nblk, device="cuda").int().view(b, nblk)
    cs = torch.tensor([2048, 100, 1025, 777], dtype=torch.int32, device="cuda")
    q = torch.randn(b, 1, h, d, device="cuda", dtype=torch.bfloat16)
    meta = fa3.get_scheduler_metadata(b, 1, nblk * page, h, hk, d, cs, page_size=page)
    assert meta.dtype == torch.int32 and int(meta[0]) >= 1 and int(meta[1]) == 8  # g = 8 heads packed for decode
    out = fa3.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cs, page_table=pt, scheduler_metadata=meta)
    ref = fa3.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cs, page_table=pt, num_splits=1)
    assert (out.float() - ref.float()).abs().max().item() < 2e-3
    with pytest.raises(RuntimeError, match="not supported"):
        fa3.flash_attn_func(q, q, q, attention_chunk=128)

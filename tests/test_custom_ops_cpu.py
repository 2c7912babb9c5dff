"""The fake (meta) implementations of the fa_b200 custom ops give the shapes / dtypes the reference's fakes give
(flash_attn/flash_attn_interface.py:117-144, 206-243, 304-338, 411-452): checked under FakeTensorMode, no GPU needed."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode


@pytest.mark.parametrize("return_softmax", [False, True])
def test_dense_fakes(return_softmax):
    import flash_attention_b200.ops  # noqa: F401  registers the ops

    with FakeTensorMode():
        q = torch.empty(2, 300, 4, 64, dtype=torch.bfloat16, device="cuda")
        k = torch.empty(2, 500, 2, 64, dtype=torch.bfloat16, device="cuda")
        out, lse, p, rng = torch.ops.fa_b200.fwd(q, k, k, 0.1, 0.125, True, -1, -1, 0.0, None, return_softmax)
        assert out.shape == q.shape and out.dtype == q.dtype
        assert lse.shape == (2, 4, 300) and lse.dtype == torch.float32
        assert p.shape == ((2, 4, 384, 512) if return_softmax else (0,)) and p.dtype == q.dtype   # seqlens rounded to 128
        assert rng.shape == (2,) and rng.dtype == torch.int64
        sd = torch.ops.fa_b200.bwd(out, q, k, k, out, lse, torch.empty_like(q), torch.empty_like(k), torch.empty_like(k), 0.1,
                                   0.125, True, -1, -1, 0.0, None, False, rng)
        assert sd.shape == (2, 4, 384) and sd.dtype == torch.float32   # softmax_d padded to 128 rows


def test_varlen_fakes():
    import flash_attention_b200.ops  # noqa: F401

    with FakeTensorMode():
        q = torch.empty(777, 8, 128, dtype=torch.float16, device="cuda")
        k = torch.empty(900, 2, 128, dtype=torch.float16, device="cuda")
        cu = torch.empty(6, dtype=torch.int32, device="cuda")
        out, lse, p, rng = torch.ops.fa_b200.varlen_fwd(q, k, k, cu, cu, 200, 333, 0.0, 0.1, False, -1, -1, 0.0, None, False, None,
                                                         None, None, False)
        assert out.shape == q.shape and lse.shape == (8, 777) and p.shape == (0,)   # unpadded LSE (h, total_q)
        sd = torch.ops.fa_b200.varlen_bwd(out, q, k, k, out, lse, torch.empty_like(q), torch.empty_like(k), torch.empty_like(k),
                                          cu, cu, 200, 333, 0.0, 0.1, False, -1, -1, 0.0, None, False, None, False)
        assert sd.shape == (8, 777 + 128 * 5)   # flash_api.cpp:1142: total_q + 128 * batch


def test_backward_schema_declares_the_in_place_outputs():
    import flash_attention_b200.ops  # noqa: F401

    schema = str(torch.ops.fa_b200.bwd.default._schema)
    for name in ("dq", "dk", "dv"):
        assert f"Tensor(a" in schema and f"!) {name}" in schema, schema   # mutates_args=("dq","dk","dv") as in the reference (:252)

"""a4: the package's own torch.compile-visible op layer.  torch.compile(fullgraph=True) must capture forward + backward
through the registered custom ops with no graph break and reproduce the eager results; torch.library.opcheck validates the
schemas (mutated dq/dk/dv), the fakes against the real kernels' outputs and the autograd registration conventions."""
import pytest
import torch

from tests.util import make_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fab():
    import flash_attention_b200 as m
    from flash_attention_b200 import _lib

    _lib.load()
    return m


@pytest.mark.parametrize("causal", [False, True])
def test_torch_compile_fullgraph_dense(fab, causal):
    q, k, v = make_qkv(2, 384, 512, 4, 2, 128)
    do = torch.randn_like(q)

    def fwd(q, k, v):
        return fab.flash_attn_func(q, k, v, causal=causal, window_size=(-1, -1))

    def step(f, q, k, v, do):   # the forward graph is captured whole; its backward is the AOT-generated graph of the same ops
        out = f(q, k, v)
        return (out,) + torch.autograd.grad(out, (q, k, v), do)

    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
    eager = step(fwd, qg, kg, vg, do)
    compiled = step(torch.compile(fwd, fullgraph=True), qg, kg, vg, do)
    assert torch.equal(eager[0], compiled[0])   # same forward kernel on the same inputs: bitwise equal
    for a, b in zip(eager[1:], compiled[1:]):   # dQ partials are reduce-added in launch order: equal up to fp32 rounding
        assert (a.float() - b.float()).abs().max().item() <= 2e-2


def test_torch_compile_fullgraph_varlen(fab):
    h, hk, d = 4, 2, 64
    lens = [100, 256, 31]
    cu = torch.tensor([0, 100, 356, 387], dtype=torch.int32, device="cuda")
    q = torch.randn(387, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(387, hk, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(387, hk, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn(387, h, d, device="cuda", dtype=torch.bfloat16)

    def fwd(q, k, v):
        return fab.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True)

    def step(f, q, k, v, do):
        out = f(q, k, v)
        return (out,) + torch.autograd.grad(out, (q, k, v), do)

    eager = step(fwd, q, k, v, do)
    compiled = step(torch.compile(fwd, fullgraph=True), q, k, v, do)
    for a, b in zip(eager, compiled):
        assert (a.float() - b.float()).abs().max().item() <= 2e-2


def test_opcheck_forward_and_backward(fab):
    from flash_attention_b200 import ops  # noqa: F401

    q, k, v = make_qkv(1, 200, 333, 4, 2, 64)
    args = (q, k, v, 0.0, 0.125, True, -1, -1, 0.0, None, False)
    torch.library.opcheck(torch.ops.fa_b200.fwd.default, args, test_utils=("test_schema", "test_faketensor"))
    out, lse, _, rng = torch.ops.fa_b200.fwd(*args)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    bargs = (torch.randn_like(out), q, k, v, out, lse, dq, dk, dv, 0.0, 0.125, True, -1, -1, 0.0, None, False, rng)
    torch.library.opcheck(torch.ops.fa_b200.bwd.default, bargs, test_utils=("test_schema", "test_faketensor"))

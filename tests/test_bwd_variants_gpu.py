"""The backward has two main-loop kernels (CTA-pair `fa_bwd2_kernel`, default for head dim 128, and the single-CTA
`fa_bwd_kernel`) and two block orders (head-major and chunked work-major).  The library picks them from
process-wide environment switches read once at first use, so each non-default combination is checked in a fresh
interpreter against the same oracle rule as tests/test_bwd_gpu.py (3x, tests/test_flash_attn.py:1130-1132)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys, torch
sys.path.insert(0, %r)
import flash_attention_b200 as fab
from tests.util import bwd_tolerances, make_qkv
for (b, sq, sk, h, hk, d, causal, window) in [(2, 300, 300, 8, 2, 128, True, (-1, -1)), (1, 384, 500, 4, 2, 128, False, (100, 37)),
                                               (1, 640, 900, 2, 2, 128, False, (-1, -1))]:
    q, k, v = make_qkv(b, sq, sk, h, hk, d)
    do = torch.randn(b, sq, h, d, device="cuda", dtype=torch.bfloat16)
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = fab.flash_attn_func(qq, kk, vv, causal=causal, window_size=window)
    out.backward(do)
    from tests.util import check_fwd
    check_fwd(out.detach(), None, q, k, v, causal=causal, window=window)
    tol = bwd_tolerances(q, k, v, do, causal=causal, window=window)
    for name, g in (("dq", qq.grad), ("dk", kk.grad), ("dv", vv.grad)):
        ref, bound = tol[name]
        err = (g.float().cpu() - ref).abs().max().item()
        assert err <= bound and not torch.isnan(g).any(), (name, err, bound)
print("VARIANT_OK")
""" % ROOT


@pytest.mark.parametrize("env", [{"FA_B200_BWD_PAIR": "0"}, {"FA_B200_BWD_PAIR": "0", "FA_B200_SCHED_CHUNK": "-1"},
                                 {"FA_B200_SCHED_CHUNK": "-4", "FA_B200_SCHED_CHUNK_FWD": "-4"},
                                 {"FA_B200_SCHED_CHUNK_FWD": "0"}])
def test_backward_kernel_variants(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "VARIANT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

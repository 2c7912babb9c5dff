"""Pin the CPU oracle against the reference's own attention_ref outputs
(fixtures written by oracle/make_golden.py from /root/reference/tests/test_flash_attn.py:217-304)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.attention_oracle import attention_oracle, attention_oracle_bwd, visible_mask

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    b, sq, sk, h, hk, d, causal, wl, wr, alibi = [int(x) for x in z["meta"]]
    t = lambda n: torch.from_numpy(z[n]).view(torch.bfloat16)
    return z, dict(b=b, sq=sq, sk=sk, h=h, hk=hk, d=d, causal=bool(causal), window=(wl, wr), alibi=bool(alibi),
                   softcap=float(z["softcap"][0]), dtype=str(z["dtype"][0])), t("q"), t("k"), t("v"), t("dout")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    z, m, q, k, v, dout = _load(path)
    slopes = torch.from_numpy(z["slopes"]) if m["alibi"] else None
    dt = torch.float32 if m["dtype"] == "fp32" else torch.bfloat16
    out, lse, dq, dk, dv = attention_oracle_bwd(
        q.to(dt), k.to(dt), v.to(dt), dout.to(dt),
        causal=m["causal"], window=m["window"], softcap=m["softcap"], alibi_slopes=slopes,
    )
    # fp32 restatement vs fp32 reference: only op-order rounding; bf16 case: reference rounds P to bf16
    tol = 2e-5 if m["dtype"] == "fp32" else 2e-2
    for name, got in (("out", out), ("dq", dq), ("dk", dk), ("dv", dv)):
        ref = torch.from_numpy(z[name])
        err = (got.float() - ref).abs().max().item()
        assert err <= tol * max(1.0, ref.abs().max().item()), (name, err)
    # lse consistency: exp(s - lse) rows sum to 1 where the reference's probabilities do
    rs = torch.from_numpy(z["attn_rowsum"])
    dead_ref = rs < 0.5
    assert torch.equal(dead_ref, torch.isinf(lse) & (lse > 0))


def test_mask_bottom_right_alignment():
    # key j visible to query i iff j <= i + sk - sq (flash_attn_interface.py:1175-1189 docstring example)
    vis = visible_mask(2, 5, causal=True)[0, 0].int().tolist()
    assert vis == [[1, 1, 1, 1, 0], [1, 1, 1, 1, 1]]
    vis = visible_mask(5, 2, causal=True)[0, 0].int().tolist()
    assert vis == [[0, 0], [0, 0], [0, 0], [1, 0], [1, 1]]


def test_dead_rows_zero_and_lse_inf():
    torch.manual_seed(1)
    q, k, v = torch.randn(1, 5, 1, 8), torch.randn(1, 2, 1, 8), torch.randn(1, 2, 1, 8)
    out, lse = attention_oracle(q, k, v, causal=True)
    assert torch.all(out[0, :3] == 0) and torch.all(torch.isinf(lse[0, 0, :3]))
    assert torch.isfinite(lse[0, 0, 3:]).all()

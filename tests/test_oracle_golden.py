"""Pin the CPU oracle against the reference's own attention_ref outputs
(fixtures written by oracle/make_golden.py from /root/reference/tests/test_flash_attn.py:217-304)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.attention_oracle import attention_oracle, attention_oracle_bwd, visible_mask

_ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLD = [p for p in _ALL if not os.path.basename(p).startswith("fp8_")]
GOLD_FP8 = [p for p in _ALL if os.path.basename(p).startswith("fp8_")]


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


def load_fp8_golden(path):
    """(meta, q8, k8, v8, descales, out_ref, out_pt, lse) of one fp8 fixture (oracle/make_golden.py:main_fp8)."""
    z = np.load(path)
    b, sq, sk, h, hk, d, causal, wl, wr, descale = [int(x) for x in z["meta"]]
    f8 = torch.float8_e4m3fn
    q8, k8, v8 = (torch.from_numpy(z[n]).view(f8) for n in ("q", "k", "v"))
    ds = [torch.from_numpy(z[n]) if descale else None for n in ("q_descale", "k_descale", "v_descale")]
    m = dict(b=b, sq=sq, sk=sk, h=h, hk=hk, d=d, causal=bool(causal), window=(wl, wr))
    return m, q8, k8, v8, ds, torch.from_numpy(z["out_ref"]), torch.from_numpy(z["out_pt"]), torch.from_numpy(z["lse"])


@pytest.mark.parametrize("path", GOLD_FP8, ids=[os.path.basename(p)[:-4] for p in GOLD_FP8])
def test_fp8_oracle_matches_reference_golden(path):
    """The fp8 branch of the oracle is pinned to the reference's FA4 oracle (flash_attn/cute/testing.py:326-470):
    fp32-upcast output + LSE, and the low-precision baseline with e4m3-rounded probabilities (`intermediate_dtype`)."""
    m, q8, k8, v8, ds, out_ref, out_pt, lse_ref = load_fp8_golden(path)
    kw = dict(causal=m["causal"], window=m["window"], q_descale=ds[0], k_descale=ds[1], v_descale=ds[2])
    bf = torch.bfloat16
    out, lse = attention_oracle(q8.to(bf).float(), k8.to(bf).float(), v8.to(bf).float(), **kw)
    # the reference returns both in the inputs' dtype (bf16): equal after the same rounding, up to one bf16 ulp where
    # fp32 op-order noise straddles a rounding boundary
    ulp = 2.0 ** -8
    assert (out.to(bf).float() - out_ref).abs().max().item() <= ulp * max(1.0, out_ref.abs().max().item())
    assert (out.to(bf).float() != out_ref).float().mean().item() < 0.01
    assert (lse.to(bf).float() - lse_ref).abs().max().item() <= ulp * max(1.0, lse_ref.abs().max().item())
    # low-precision baseline: bf16 math, K scaled before the GEMM, softmax(S) rounded to e4m3
    pt, _ = attention_oracle(q8.to(bf), k8.to(bf), v8.to(bf), compute_dtype=bf, reorder_ops=True,
                             p_dtype=torch.float8_e4m3fn, p_quant="normalized", **kw)
    err_pt = (pt.float() - out_pt).abs().max().item()
    base = (out_pt - out_ref).abs().max().item()   # what the rounding itself costs (the tolerance yardstick)
    assert err_pt <= 0.5 * base + 1e-3, (err_pt, base)
    # and the yardstick itself agrees: |pt - ref| of the restatement within 25 % of the reference's
    assert abs((pt.float() - out).abs().max().item() - base) <= 0.25 * base + 1e-3


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

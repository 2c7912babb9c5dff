"""Shared helpers for the parity tests (CUDA path vs the CPU oracle)."""
import math

import torch

from oracle.attention_oracle import attention_oracle, attention_oracle_bwd  # checker only


def make_qkv(b, sq, sk, h, hk, d, dtype=torch.bfloat16, seed=0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(b, sq, h, d, generator=g).to(dtype)
    k = torch.randn(b, sk, hk, d, generator=g).to(dtype)
    v = torch.randn(b, sk, hk, d, generator=g).to(dtype)
    return q.to(device), k.to(device), v.to(device)


def fwd_tolerances(q, k, v, **kw):
    """The reference's parity rule (tests/test_flash_attn.py:1121, tests/cute/test_flash_attn.py:462-463):
    |out - ref_fp32| <= 2 * |ref_lowp - ref_fp32| + atol, atol = 2 * |ref + 0.3 - 0.3 - ref|."""
    qc, kc, vc = q.cpu(), k.cpu(), v.cpu()
    kw = {k_: (v_.cpu() if torch.is_tensor(v_) else v_) for k_, v_ in kw.items()}
    ref, lse = attention_oracle(qc, kc, vc, **kw)  # fp32 math, output rounded to q.dtype
    ref32, _ = attention_oracle(qc.float(), kc.float(), vc.float(), **kw)
    lowp, _ = attention_oracle(qc, kc, vc, compute_dtype=q.dtype, **kw)
    atol = 2 * (ref32 + 0.3 - 0.3 - ref32).abs().max().item()
    bound = 2 * (lowp.float() - ref32).abs().max().item() + atol
    return ref32, lse, bound


def check_fwd(out, lse, q, k, v, name="", **kw):
    ref32, lse_ref, bound = fwd_tolerances(q, k, v, **kw)
    err = (out.float().cpu() - ref32).abs().max().item()
    assert not torch.isnan(out).any(), f"{name}: NaN in out"
    assert err <= bound, f"{name}: out max err {err:.3e} > bound {bound:.3e}"
    if lse is not None:
        lse_c = lse.float().cpu()
        fin = torch.isfinite(lse_ref)
        assert torch.equal(torch.isfinite(lse_c), fin), f"{name}: lse finiteness pattern differs"
        lerr = (lse_c[fin] - lse_ref[fin]).abs().max().item() if fin.any() else 0.0
        assert lerr <= 2e-3 * max(1.0, lse_ref[fin].abs().max().item() if fin.any() else 1.0), f"{name}: lse err {lerr:.3e}"
        if (~fin).any():
            assert torch.equal(lse_c[~fin], lse_ref[~fin]), f"{name}: lse inf sign differs"
    return err, bound


def bwd_tolerances(q, k, v, dout, **kw):
    """dq/dk/dv: |g - ref_fp32| <= 3 * |ref_lowp - ref_fp32| + atol (tests/test_flash_attn.py:1130-1132)."""
    qc, kc, vc, dc = q.cpu(), k.cpu(), v.cpu(), dout.cpu()
    kw = {k_: (v_.cpu() if torch.is_tensor(v_) else v_) for k_, v_ in kw.items()}
    _, _, dq, dk, dv = attention_oracle_bwd(qc, kc, vc, dc, **kw)
    # low-precision reference: same math with bf16/fp16 intermediates (autograd through the lowp forward)
    ql = qc.clone().requires_grad_(True)
    kl = kc.clone().requires_grad_(True)
    vl = vc.clone().requires_grad_(True)
    out_l, _ = attention_oracle(ql, kl, vl, compute_dtype=q.dtype, **kw)
    dql, dkl, dvl = torch.autograd.grad(out_l, (ql, kl, vl), dc)
    res = {}
    for n, g32, gl in (("dq", dq, dql), ("dk", dk, dkl), ("dv", dv, dvl)):
        atol = 2 * (g32 + 0.3 - 0.3 - g32).abs().max().item()
        res[n] = (g32, 3 * (gl.float() - g32).abs().max().item() + atol)
    return res


_REAL_FA2 = None


def load_reference_fa2():
    """The reference's own compiled FA2 kernels: the `flash_attn_2_cuda` extension of the flash_attn wheel installed in the
    image (sm80 mma.sync code built for sm_100).  Imported once per process under its real name even if
    `flash_attention_b200.install()` has since put this library's operator module into sys.modules under that name."""
    global _REAL_FA2
    if _REAL_FA2 is None:
        import importlib
        import sys

        import pytest

        ours = sys.modules.pop("flash_attn_2_cuda", None)
        try:
            mod = importlib.import_module("flash_attn_2_cuda")
        except Exception as e:  # noqa: BLE001
            pytest.skip(f"reference FA2 extension not importable: {e}")
        finally:
            if ours is not None:
                sys.modules["flash_attn_2_cuda"] = ours
            else:
                sys.modules.pop("flash_attn_2_cuda", None)
        if not getattr(mod, "__file__", "").endswith(".so"):
            pytest.skip("flash_attn_2_cuda resolves to a non-compiled module")
        # it must really be the reference's binary, not this package's module under the same name (pybind11 >= 3 caches
        # extension modules by spec name: see flash_attention_b200/flash_attn_2_cuda.py)
        assert not hasattr(mod, "fwd_fp8") and "flash_attention_b200" not in mod.__file__, \
            f"flash_attn_2_cuda resolved to this package's own module: {mod.__file__}"
        _REAL_FA2 = mod
    return _REAL_FA2

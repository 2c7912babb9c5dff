"""Generate tests/golden/*.npz from the REFERENCE's own attention_ref.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py

The reference test module cannot be imported without a GPU (it queries the device
capability at import, tests/test_flash_attn.py:23-26), so the three pure-torch
functions (attn_bias_from_alibi_slopes :29, construct_local_mask :182,
attention_ref :217) are pulled out of its AST and executed unmodified on CPU.
Nothing is copied into this repo: only their numeric outputs are stored.
"""
import ast
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference/tests/test_flash_attn.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
WANTED = {"attn_bias_from_alibi_slopes", "construct_local_mask", "attention_ref"}


def load_reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    mod = ast.Module(body=body, type_ignores=[])
    ns = {}
    exec("import math, torch\nfrom einops import rearrange, repeat\n", ns)
    exec(compile(mod, REF, "exec"), ns)
    return ns


CASES = [
    # name, b, sq, sk, h, hk, d, causal, window, softcap, alibi, dtype
    ("c1_noncausal", 2, 128, 128, 4, 4, 64, False, (-1, -1), 0.0, False, "fp32"),
    ("c1_causal", 2, 128, 128, 4, 4, 64, True, (-1, -1), 0.0, False, "fp32"),
    ("gqa_causal_sq_ne_sk", 1, 96, 160, 4, 2, 64, True, (-1, -1), 0.0, False, "fp32"),
    ("local_window", 1, 128, 144, 2, 2, 32, False, (17, 9), 0.0, False, "fp32"),
    ("alibi_noncausal", 2, 64, 80, 4, 4, 32, False, (-1, -1), 0.0, True, "fp32"),
    ("alibi_causal", 2, 64, 80, 4, 4, 32, True, (-1, -1), 0.0, True, "fp32"),
    ("softcap", 1, 64, 64, 2, 1, 64, False, (-1, -1), 15.0, False, "fp32"),
    ("bf16_hd128_causal", 1, 192, 192, 2, 1, 128, True, (-1, -1), 0.0, False, "bf16"),
    ("sq_gt_sk_causal_deadrows", 1, 64, 32, 2, 2, 32, True, (-1, -1), 0.0, False, "fp32"),
]


def main():
    ns = load_reference_functions()
    os.makedirs(OUT, exist_ok=True)
    for name, b, sq, sk, h, hk, d, causal, window, softcap, alibi, dt in CASES:
        torch.manual_seed(0)
        dtype = torch.float32 if dt == "fp32" else torch.bfloat16
        # inputs are bf16-representable so that they can be stored as 16-bit patterns
        q = torch.randn(b, sq, h, d).bfloat16()
        k = torch.randn(b, sk, hk, d).bfloat16()
        v = torch.randn(b, sk, hk, d).bfloat16()
        dout = torch.randn(b, sq, h, d).bfloat16()
        slopes = (torch.rand(b, h) * 0.3).float() if alibi else None
        bias = None
        if alibi:
            bias = ns["attn_bias_from_alibi_slopes"](slopes, sq, sk, causal=causal)
        qq = q.to(dtype).requires_grad_(True)
        kk = k.to(dtype).requires_grad_(True)
        vv = v.to(dtype).requires_grad_(True)
        out, attn = ns["attention_ref"](
            qq, kk, vv, None, None, bias, 0.0, None, causal=causal, window_size=window, softcap=softcap
        )
        dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), dout.to(dtype))
        # lse from the reference's own probabilities is not exposed; store row-max prob instead
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            q=q.view(torch.int16).numpy(),
            k=k.view(torch.int16).numpy(),
            v=v.view(torch.int16).numpy(),
            dout=dout.view(torch.int16).numpy(),
            slopes=(slopes.numpy() if slopes is not None else np.zeros(0, np.float32)),
            out=out.detach().float().numpy(),
            dq=dq.float().numpy(),
            dk=dk.float().numpy(),
            dv=dv.float().numpy(),
            attn_rowsum=attn.detach().float().sum(-1).numpy(),
            meta=np.array([b, sq, sk, h, hk, d, int(causal), window[0], window[1], int(alibi)], np.int64),
            softcap=np.array([softcap], np.float64),
            dtype=np.array([dt]),
        )
        print("wrote", name, tuple(out.shape), float(out.float().abs().max()))


FP8_REF = "/root/reference/flash_attn/cute/testing.py"
FP8_CASES = [
    # name, b, sq, sk, h, hk, d, causal, window(None = off), descales
    ("fp8_noncausal_descale", 2, 96, 160, 4, 2, 128, False, (None, None), True),
    ("fp8_causal", 1, 128, 128, 2, 2, 128, True, (None, None), False),
    ("fp8_window_c4_mask", 1, 192, 192, 2, 1, 128, True, (48, 0), True),   # BASELINE C4's mask family (causal sliding window)
    ("fp8_hd64_gqa", 1, 80, 112, 4, 1, 64, False, (None, None), True),
]


def main_fp8():
    """FP8 e4m3 forward fixtures from the reference's FA4 oracle (flash_attn/cute/testing.py:326-470), loaded by file
    path (pure torch, runs on CPU).  Stored per case, exactly as the reference's fp8 tests compute them
    (tests/cute/test_flash_attn.py:413-448): `out_ref` = fp32 upcast, `out_pt` = bf16 math with reordered ops and the
    softmax probabilities rounded to e4m3 (`intermediate_dtype`), plus the LSE of the upcast run."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_cute_testing", FP8_REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    f8 = torch.float8_e4m3fn
    for name, b, sq, sk, h, hk, d, causal, window, descale in FP8_CASES:
        torch.manual_seed(0)
        # e4m3-representable inputs held in bf16 (the reference's dtype_ref for fp8), stored as raw e4m3 bytes
        q8, k8, v8 = (torch.randn(b, s_, h_, d).to(f8) for s_, h_ in ((sq, h), (sk, hk), (sk, hk)))
        q, k, v = (t.to(torch.bfloat16) for t in (q8, k8, v8))
        ds = [(torch.rand(b, hk) + 0.5).float() if descale else None for _ in range(3)]
        kw = dict(causal=causal, q_descale=ds[0], k_descale=ds[1], v_descale=ds[2], window_size=window)
        out_ref, _, lse = ref.attention_ref(q, k, v, None, None, return_lse=True, **kw)
        out_pt, _ = ref.attention_ref(q, k, v, None, None, upcast=False, reorder_ops=True, intermediate_dtype=f8, **kw)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            q=q8.view(torch.uint8).numpy(), k=k8.view(torch.uint8).numpy(), v=v8.view(torch.uint8).numpy(),
            q_descale=(ds[0].numpy() if descale else np.zeros(0, np.float32)),
            k_descale=(ds[1].numpy() if descale else np.zeros(0, np.float32)),
            v_descale=(ds[2].numpy() if descale else np.zeros(0, np.float32)),
            out_ref=out_ref.float().numpy(), out_pt=out_pt.float().numpy(), lse=lse.float().numpy(),
            meta=np.array([b, sq, sk, h, hk, d, int(causal), -1 if window[0] is None else window[0],
                           -1 if window[1] is None else window[1], int(descale)], np.int64),
        )
        print("wrote", name, tuple(out_ref.shape), float((out_pt.float() - out_ref.float()).abs().max()))


if __name__ == "__main__":
    main()
    main_fp8()

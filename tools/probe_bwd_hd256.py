"""Parity probe of the head-dim 129..256 backward (fa_bwd256_kernel, opt-in: FA_B200_BWD_HD256=1) against the CPU oracle
with the reference's 3x rule.  Run on the GPU box:  FA_B200_BWD_HD256=1 python tools/probe_bwd_hd256.py"""
import os
import sys
import time

os.environ.setdefault("FA_B200_BWD_HD256", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import flash_attention_b200 as fab
from tests.util import bwd_tolerances, make_qkv

CASES = [
    # b, sq, sk, h, hk, d, causal, window, softcap, dtype, deterministic
    (1, 64, 128, 1, 1, 256, False, (-1, -1), 0.0, torch.bfloat16, False),
    (1, 128, 128, 1, 1, 256, False, (-1, -1), 0.0, torch.bfloat16, False),
    (2, 256, 384, 2, 2, 256, False, (-1, -1), 0.0, torch.bfloat16, False),
    (2, 113, 203, 3, 3, 256, True, (-1, -1), 0.0, torch.bfloat16, False),
    (2, 113, 203, 3, 3, 192, False, (-1, -1), 0.0, torch.float16, False),
    (1, 1023, 1024, 2, 1, 192, True, (-1, -1), 0.0, torch.bfloat16, False),
    (1, 512, 256, 2, 2, 256, True, (-1, -1), 0.0, torch.bfloat16, False),
    (1, 384, 500, 4, 2, 224, False, (100, 37), 0.0, torch.bfloat16, False),
    (1, 700, 700, 1, 1, 160, True, (129, -1), 0.0, torch.bfloat16, False),
    (1, 256, 256, 2, 2, 256, False, (-1, -1), 15.0, torch.bfloat16, False),
    (2, 300, 300, 8, 2, 256, True, (-1, -1), 0.0, torch.bfloat16, True),
    (1, 1024, 1024, 4, 4, 256, True, (-1, -1), 0.0, torch.bfloat16, True),
]


def grads(q, k, v, do, causal, window, softcap, det):
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = fab.flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, deterministic=det)
    out.backward(do)
    return q.grad, k.grad, v.grad


def main():
    bad = 0
    for (b, sq, sk, h, hk, d, causal, window, softcap, dtype, det) in CASES:
        q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype)
        do = torch.randn(b, sq, h, d, device="cuda", dtype=dtype)
        try:
            dq, dk, dv = grads(q, k, v, do, causal, window, softcap, det)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"FAIL {b, sq, sk, h, hk, d, causal, window, softcap, det}: {type(e).__name__}: {e}", flush=True)
            bad += 1
            if "CUDA" in str(e) or "launch" in str(e):
                break
            continue
        tol = bwd_tolerances(q, k, v, do, causal=causal, window=window, softcap=softcap)
        msg = []
        ok = True
        for name, g in (("dq", dq), ("dk", dk), ("dv", dv)):
            ref, bound = tol[name]
            err = (g.float().cpu() - ref).abs().max().item()
            nan = bool(torch.isnan(g).any())
            ok = ok and (not nan) and err <= bound
            msg.append(f"{name} {err:.2e}/{bound:.2e}{' NaN' if nan else ''}")
        if det and ok:
            dq2, dk2, dv2 = grads(q, k, v, do, causal, window, softcap, det)
            same = torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)
            ok = ok and same
            msg.append("bitwise " + ("same" if same else "DIFFERENT"))
        bad += 0 if ok else 1
        print(("ok   " if ok else "FAIL ") + f"{b, sq, sk, h, hk, d, causal, window, softcap, det}: " + ", ".join(msg), flush=True)
    # speed at the reference's hdim-256 benchmark shape (benchmark_flash_attention.py: 16k tokens, dim 2048 -> 8 heads)
    if bad == 0:
        for causal in (False, True):
            b, s, h, d = 2, 8192, 8, 256
            q, k, v = make_qkv(b, s, s, h, h, d, torch.bfloat16)
            q, k, v = (t.requires_grad_(True) for t in (q, k, v))
            do = torch.randn_like(q)
            out = fab.flash_attn_func(q, k, v, causal=causal)
            for _ in range(2):
                out.backward(do, retain_graph=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 5
            for _ in range(n):
                out.backward(do, retain_graph=True)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / n * 1e3
            fl = 2.5 * 4 * b * h * s * s * d / (2 if causal else 1)
            print(f"bwd b{b} h{h} s{s} d{d} causal={causal}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

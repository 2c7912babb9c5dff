"""Time the reference's own kernels on this B200 (context for BASELINE.md): installed FA2 2.8.3 wheel
(sm80 mma.sync code recompiled for sm_100) and, if it runs with the image's cutlass-dsl, reference FA4
(flash_attn/cute installed unmodified into baseline/_ref).  usage: python tools/baseline_probe.py fa2|fa4"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench(f, n=10, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def cases():
    yield "C2 fwd noncausal b4 h16 s8192", 4, 16, 16, 8192, False
    yield "headline causal b4 h16 s8192", 4, 16, 16, 8192, True
    yield "C3 causal gqa b2 h32/4 s16384", 2, 32, 4, 16384, True


def run(fn_name, flash_attn_func):
    for name, b, h, hk, s, causal in cases():
        torch.manual_seed(0)
        q = torch.randn(b, s, h, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(b, s, hk, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(b, s, hk, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        do = torch.randn(b, s, h, 128, device="cuda", dtype=torch.bfloat16)
        fl = 4 * b * h * s * s * 128 / (2 if causal else 1)
        try:
            def fwd():
                o = flash_attn_func(q, k, v, causal=causal)
                return o[0] if isinstance(o, tuple) else o
            ms = bench(lambda: fwd())
            print(f"[{fn_name}] {name} fwd {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
            def fb():
                o = fwd()
                o.backward(do)
            ms2 = bench(fb, n=5)
            print(f"[{fn_name}] {name} fwd+bwd {ms2:.3f} ms {3.5 * fl / ms2 / 1e9:.1f} TFLOP/s (bwd-only {(ms2 - ms):.3f} ms {2.5 * fl / (ms2 - ms) / 1e9:.1f})", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[{fn_name}] {name} FAILED: {type(e).__name__}: {str(e)[:500]}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "fa2":
        from flash_attn import flash_attn_func
        run("FA2-2.8.3", flash_attn_func)
        # decode C5 (scaled to fit quickly): paged kv
        from flash_attn import flash_attn_with_kvcache
        b, h, hk, d, page, kv = 64, 32, 4, 128, 256, 65536
        nb = kv // page
        kc = torch.randn(b * nb, page, hk, d, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(b * nb, page, hk, d, device="cuda", dtype=torch.bfloat16)
        bt = torch.randperm(b * nb, device="cuda").int().view(b, nb)
        cs = torch.full((b,), kv, dtype=torch.int32, device="cuda")
        q = torch.randn(b, 1, h, d, device="cuda", dtype=torch.bfloat16)
        ms = bench(lambda: flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cs, block_table=bt), n=5)
        byts = 2 * b * kv * hk * d * 2
        print(f"[FA2-2.8.3] C5 decode paged {ms:.3f} ms {byts / ms / 1e6:.1f} GB/s", flush=True)
    elif which == "fa4":
        import flash_attn
        flash_attn.__path__.append(os.path.join(ROOT, "baseline", "_ref", "flash_attn"))
        t0 = time.time()
        from flash_attn.cute import flash_attn_func
        print("imported flash_attn.cute in %.1fs" % (time.time() - t0), flush=True)
        run("FA4-ref", flash_attn_func)

"""Same-box comparators for BASELINE.md: time the reference's own kernels and the vendor's on this B200, on the BASELINE
shapes, with CUDA events (3 warm-up + 10 timed, inputs 4 x 128 MiB > L2).

  python tools/baseline_probe.py fa2     installed FA2 2.8.3 wheel (the reference's sm80 mma.sync kernels built for sm_100)
  python tools/baseline_probe.py fa4     the reference's Blackwell kernels (flash_attn/cute, UNMODIFIED, from baseline/_ref)
                                         traced through tools/fa4_shim.py (the image's cutlass-dsl is older than FA4 asks for)
  python tools/baseline_probe.py cudnn   cuDNN SDPA through the frontend graph API (reference harness:
                                         flash_attn/cute/bench_utils.py:126-243), fwd and bwd
  python tools/baseline_probe.py sdpa    torch SDPA pinned to its cuDNN backend (cross-check of the line above)
  python tools/baseline_probe.py ours    this library, same loop (so that every row of the table is one lease)

Each line is also appended as JSON to gpurun_out/r02_baseline.jsonl."""
import json
import math
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "r02_baseline.jsonl")


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
    yield "C2/headline noncausal b4 h16 s8192", 4, 16, 16, 8192, False
    yield "headline causal b4 h16 s8192", 4, 16, 16, 8192, True
    yield "C3 causal gqa b2 h32/4 s16384", 2, 32, 4, 16384, True
    yield "short causal b32 h16 s1024", 32, 16, 16, 1024, True
    yield "short causal b16 h16 s2048", 16, 16, 16, 2048, True


def emit(impl, name, what, ms, fl):
    rec = {"impl": impl, "case": name, "what": what, "ms": ms, "tflops": fl / ms / 1e9}
    print(f"[{impl}] {name} {what} {ms:.3f} ms {rec['tflops']:.1f} TFLOP/s", flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")


def make(b, h, hk, s, d=128, grad=True):
    torch.manual_seed(0)
    q = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=grad)
    k = torch.randn(b, s, hk, d, device="cuda", dtype=torch.bfloat16, requires_grad=grad)
    v = torch.randn(b, s, hk, d, device="cuda", dtype=torch.bfloat16, requires_grad=grad)
    do = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16)
    return q, k, v, do


def run_autograd(impl, flash_attn_func):
    """fwd alone, then fwd+bwd through autograd; bwd = difference (what a user of the public API sees)."""
    for name, b, h, hk, s, causal in cases():
        q, k, v, do = make(b, h, hk, s)
        fl = 4 * b * h * s * s * 128 / (2 if causal else 1)
        try:
            def fwd():
                o = flash_attn_func(q, k, v, causal=causal)
                return o[0] if isinstance(o, tuple) else o
            ms = bench(fwd)
            emit(impl, name, "fwd", ms, fl)

            def fb():
                fwd().backward(do)
                q.grad = k.grad = v.grad = None
            ms2 = bench(fb, n=5)
            emit(impl, name, "fwd+bwd", ms2, 3.5 * fl)
            emit(impl, name, "bwd(diff)", ms2 - ms, 2.5 * fl)
        except Exception as e:  # noqa: BLE001
            print(f"[{impl}] {name} FAILED: {type(e).__name__}: {str(e)[:800]}", flush=True)
            traceback.print_exc(limit=6)


def run_cudnn():
    from tools import fa4_shim  # only for the path to baseline/_ref (the harness lives in the reference tree)

    fa4_shim.apply()
    from flash_attn.cute import bench_utils as bu

    for name, b, h, hk, s, causal in cases():
        q, k, v, do = make(b, h, hk, s, grad=False)
        fl = 4 * b * h * s * s * 128 / (2 if causal else 1)
        try:
            # cuDNN layout (b, h, s, d) as strided views of the same (b, s, h, d) memory (what the reference harness does)
            qt, kt, vt, gt = (t.transpose(1, 2) for t in (q, k, v, do))
            fwd_fn, o_gpu, stats = bu.cudnn_fwd_setup(qt, kt, vt, causal=causal)
            ms = bench(fwd_fn)
            emit("cuDNN-9.22-frontend-1.18", name, "fwd", ms, fl)
            bwd_fn = bu.cudnn_bwd_setup(qt, kt, vt, o_gpu, gt, stats, causal=causal)
            msb = bench(bwd_fn, n=5)
            emit("cuDNN-9.22-frontend-1.18", name, "bwd", msb, 2.5 * fl)
            emit("cuDNN-9.22-frontend-1.18", name, "fwd+bwd", ms + msb, 3.5 * fl)
        except Exception as e:  # noqa: BLE001
            print(f"[cudnn] {name} FAILED: {type(e).__name__}: {str(e)[:800]}", flush=True)
            traceback.print_exc(limit=6)


def run_sdpa():
    from torch.nn.attention import SDPBackend, sdpa_kernel
    import torch.nn.functional as F

    def f(q, k, v, causal=False):
        g = q.shape[2] // k.shape[2]
        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal,
                                                  enable_gqa=g > 1).transpose(1, 2)
    run_autograd("torch-sdpa-cudnn", f)


def run_fa4():
    from tools import fa4_shim

    print("shim:", fa4_shim.apply(), flush=True)
    t0 = time.time()
    from flash_attn.cute import flash_attn_func
    print("imported flash_attn.cute in %.1fs" % (time.time() - t0), flush=True)
    run_autograd("FA4-ref(cute,shim)", flash_attn_func)
    # C4: fp8 e4m3 forward, s32768 window (4096, 0)
    try:
        b, h, s, d, w = 1, 16, 32768, 128, 4096
        f8 = torch.float8_e4m3fn
        q, k, v, _ = make(b, h, h, s, grad=False)
        pairs = sum(min(i, w) + 1 for i in range(s))
        fl = 4.0 * b * h * pairs * d
        for dt_name, cast in (("fp8", lambda t: t.to(f8)), ("bf16", lambda t: t)):
            qq, kk, vv = cast(q), cast(k), cast(v)
            ms = bench(lambda: flash_attn_func(qq, kk, vv, causal=True, window_size=(w, 0)))
            emit("FA4-ref(cute,shim)", f"C4 {dt_name} s32768 window(4096,0)", "fwd", ms, fl)
    except Exception as e:  # noqa: BLE001
        print(f"[FA4] C4 FAILED: {type(e).__name__}: {str(e)[:800]}", flush=True)
        traceback.print_exc(limit=6)


def run_ours():
    import flash_attention_b200 as fab

    run_autograd("fa_b200", fab.flash_attn_func)
    b, h, s, d, w = 1, 16, 32768, 128, 4096
    f8 = torch.float8_e4m3fn
    q, k, v, _ = make(b, h, h, s, grad=False)
    pairs = sum(min(i, w) + 1 for i in range(s))
    fl = 4.0 * b * h * pairs * d
    for dt_name, cast in (("fp8", lambda t: t.to(f8)), ("bf16", lambda t: t)):
        qq, kk, vv = cast(q), cast(k), cast(v)
        ms = bench(lambda: fab.flash_attn_func(qq, kk, vv, causal=True, window_size=(w, 0)))
        emit("fa_b200", f"C4 {dt_name} s32768 window(4096,0)", "fwd", ms, fl)


def run_fa2():
    from flash_attn import flash_attn_func, flash_attn_with_kvcache

    run_autograd("FA2-2.8.3", flash_attn_func)
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


if __name__ == "__main__":
    {"fa2": run_fa2, "fa4": run_fa4, "cudnn": run_cudnn, "sdpa": run_sdpa, "ours": run_ours}[sys.argv[1]]()

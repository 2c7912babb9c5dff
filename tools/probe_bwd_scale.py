"""Is the backward limited per SM (TMA reduce engine) or chip-wide (L2 atomics)?  Time the bwd kernel alone at
grids that occupy 64 / 128 / 148+ SMs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_b200 import flash_attn_2_cuda as op
for (b, h, s) in ((1, 1, 8192), (1, 2, 8192), (1, 4, 8192), (4, 16, 8192)):
    q = torch.randn(b, s, h, 128, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
    scale = 128 ** -0.5
    out, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, scale, False, -1, -1, 0.0, False, None)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    f = lambda: op.bwd(do, q, k, v, out, lse, dq, dk, dv, None, 0.0, scale, False, -1, -1, 0.0, False, None, None)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ctas = (s // 128) * h * b
    iters = s // 128
    print(f"[bwd-scale] b{b} h{h} s{s}: {ms:.3f} ms, {ctas} CTAs, per-CTA-iteration {ms * 1e3 / (iters * max(1, -(-ctas // 148))):.2f} us, {2.5 * 4 * b * h * s * s * 128 / ms / 1e9:.0f} TFLOP/s", flush=True)

"""Timing of the wide-S forward kernel vs the two-tile kernel on the BASELINE forward shapes (one line per config)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa  # noqa: E402

tag = os.environ.get("FA_TAG", "default")


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


torch.manual_seed(0)
out = []
for name, b, h, hk, s, d, causal in (("C2", 4, 16, 16, 8192, 128, False), ("causal8k", 4, 16, 16, 8192, 128, True),
                                      ("C3fwd", 2, 32, 4, 16384, 128, True), ("nc2k", 16, 16, 16, 2048, 128, False),
                                      ("causal2k", 16, 16, 16, 2048, 128, True)):
    q = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(b, s, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(b, s, hk, d, device="cuda", dtype=torch.bfloat16)
    ms = bench(lambda: fa.flash_attn_func(q, k, v, causal=causal))
    fl = 4 * b * h * s * s * d / (2 if causal else 1)
    out.append(f"{name} {fl / ms / 1e9:.0f}")
print(f"[{tag}] " + "  ".join(out), flush=True)

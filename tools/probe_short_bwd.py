"""Backward throughput at short / medium sequence lengths (32k tokens per batch, hidden 2048, hd128), the operator module's `bwd`
timed alone (preprocess + main kernel + dQ convert)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa
from flash_attention_b200 import flash_attn_2_cuda as op

def timeit(f, n=10, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

d, h = 128, 16
out = []
for causal in (True, False):
    for s in (512, 1024, 2048, 4096):
        b = 32768 // s
        q, k, v, do = (torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
        o, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, d ** -0.5, causal, -1, -1, 0.0, False, None)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ms = timeit(lambda: op.bwd(do, q, k, v, o, lse, dq, dk, dv, None, 0.0, d ** -0.5, causal, -1, -1, 0.0, False, None, None))
        fl = 2.5 * 4 * b * h * s * s * d / (2 if causal else 1)
        out.append(f"{'c' if causal else 'n'}{s} {fl / ms / 1e9:.0f}")
print("  ".join(out), flush=True)

"""Forward / backward throughput at short sequence lengths (32k tokens per batch, hidden 2048, hd128)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa

def timeit(f, n=10, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

d, h = 128, 16
for causal in (True, False):
    for s in (512, 1024, 2048, 4096):
        b = 32768 // s
        q, k, v = (torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
        fl = 4 * b * h * s * s * d / (2 if causal else 1)
        ms = timeit(lambda: fa.flash_attn_func(q, k, v, causal=causal))
        print(f"causal={causal} s={s}: fwd {ms:.4f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa
for (b, s, h, d, causal) in ((4, 8192, 32, 64, False), (4, 8192, 32, 64, True), (4, 8192, 16, 128, False)):
    q = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
    f = lambda: fa.flash_attn_func(q, k, v, causal=causal)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"[d64 probe] d{d} causal={causal}: {ms:.3f} ms {4*b*h*s*s*d/(2 if causal else 1)/ms/1e9:.1f} TFLOP/s", flush=True)

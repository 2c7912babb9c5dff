"""Host-side cost per call of the operator module (tiny problem, so the GPU is never the bottleneck)."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa

q, k, v = (torch.randn(1, 128, 2, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
do = torch.randn_like(q)

def bench(name, f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{name}: {dt * 1e6:.1f} us per call", flush=True)

def fb():
    o = fa.flash_attn_func(qg, kg, vg, causal=True); o.backward(do); qg.grad = kg.grad = vg.grad = None

bench("forward (no grad)", lambda: fa.flash_attn_func(q, k, v, causal=True))
bench("forward + backward (autograd)", fb, 1000)
pr = cProfile.Profile(); pr.enable()
for _ in range(500): fa.flash_attn_func(q, k, v, causal=True)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

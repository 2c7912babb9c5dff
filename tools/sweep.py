"""Sweep in the style of the reference's benchmark (assets/fa4_paper.pdf sect. 5: batch*seqlen = 32k tokens, hidden 2048,
hd 64/128, causal and not): forward and forward+backward TFLOP/s.  usage: python tools/sweep.py [ours|fa2]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "ours"
if which == "ours":
    import flash_attention_b200 as fa
else:
    import flash_attn as fa

def timeit(f, n=10, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for d in (128, 64):
    h = 2048 // d
    for causal in (False, True):
        for s in (1024, 2048, 4096, 8192, 16384, 32768):
            b = 32768 // s
            q = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
            k = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
            v = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
            do = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16)
            fl = 4 * b * h * s * s * d / (2 if causal else 1)
            fwd = lambda: fa.flash_attn_func(q, k, v, causal=causal)
            def fb():
                o = fwd(); o.backward(do); q.grad = k.grad = v.grad = None
            mf = timeit(fwd); mb = timeit(fb, n=5)
            print(json.dumps(dict(impl=which, d=d, causal=causal, seqlen=s, batch=b, fwd_ms=round(mf, 4), fwd_tflops=round(fl / mf / 1e9, 1),
                                  fwdbwd_ms=round(mb, 4), fwdbwd_tflops=round(3.5 * fl / mb / 1e9, 1),
                                  bwd_tflops=round(2.5 * fl / (mb - mf) / 1e9, 1))), flush=True)
            del q, k, v, do

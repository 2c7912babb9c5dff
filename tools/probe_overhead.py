"""Host-side cost per call (tiny problem, so the GPU is never the bottleneck): the compiled operator module's `fwd`
directly (what the reference's interface calls), this package's public API (eager fast path), and forward + backward
through autograd.  VERDICT r1 target: <= 10 us per forward call of the operator module (was 44 us with the ctypes shim)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa  # noqa: E402
from flash_attention_b200 import flash_attn_2_cuda as op  # noqa: E402

q, k, v = (torch.randn(1, 128, 2, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
do = torch.randn_like(q)
scale = 128 ** -0.5


def bench(name, f, n=3000):
    for _ in range(100):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    t1 = time.perf_counter()   # host time to ENQUEUE n calls
    torch.cuda.synchronize()
    t2 = time.perf_counter()   # ... and until the GPU has finished them (if this is much larger the launch queue was full
    #                            and the enqueue figure is a GPU-bound number, not host overhead)
    print(f"{name}: {(t1 - t0) / n * 1e6:.1f} us per call host enqueue, {(t2 - t0) / n * 1e6:.1f} us incl. GPU drain", flush=True)


def fb():
    o = fa.flash_attn_func(qg, kg, vg, causal=True)
    o.backward(do)
    qg.grad = kg.grad = vg.grad = None


out = torch.empty_like(q)
bench("operator module fwd (out_ given)", lambda: op.fwd(q, k, v, out, None, 0.0, scale, True, -1, -1, 0.0, False, None))
bench("operator module fwd (allocates out)", lambda: op.fwd(q, k, v, None, None, 0.0, scale, True, -1, -1, 0.0, False, None))
with torch.no_grad():
    bench("public flash_attn_func (no grad)", lambda: fa.flash_attn_func(q, k, v, causal=True))
bench("public flash_attn_func forward + backward (autograd)", fb, 1500)

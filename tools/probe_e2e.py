"""e2e (host buffers) step time of the host pipeline for different piece counts, plus raw PCIe copy rates."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_b200.host_pipeline import HostAttentionPipeline

b, s, h, d = 4, 8192, 16, 128
pin = lambda: torch.randn(b, s, h, d).bfloat16().pin_memory()
hq, hk, hv, hdo = pin(), pin(), pin(), pin()
ho, hdq, hdk, hdv = (torch.empty_like(hq).pin_memory() for _ in range(4))
dev_t = torch.empty(b, s, h, d, dtype=torch.bfloat16, device="cuda")


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(lambda: dev_t.copy_(hq, non_blocking=True))
print(f"linear H2D 128 MiB: {ms:.3f} ms = {hq.numel() * 2 / ms / 1e6:.1f} GB/s")
ms = timed(lambda: ho.copy_(dev_t, non_blocking=True))
print(f"linear D2H 128 MiB: {ms:.3f} ms = {hq.numel() * 2 / ms / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        dev_t.copy_(hq, non_blocking=True)
    with torch.cuda.stream(s2):
        ho.copy_(dev_t, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
ms = timed(both)
print(f"H2D + D2H concurrently, 128 MiB each: {ms:.3f} ms = {hq.numel() * 2 / ms / 1e6:.1f} GB/s per direction")
fl = 3.5 * 4 * b * h * s * s * d
for chunks, hg in ((4, 1), (4, 2), (4, 4), (4, 8), (4, 16), (2, 16), (1, 16)):
    pipe = HostAttentionPipeline(b, s, s, h, h, d, torch.bfloat16, "cuda", chunks=chunks, backward=True, head_groups=hg)
    ms = timed(lambda: pipe.run(hq, hk, hv, hdo, ho, hdq, hdk, hdv, causal=False))
    print(f"pieces {chunks * hg:3d} (batch chunks {chunks}, head groups {hg:2d}): {ms:.3f} ms/step  {fl / ms / 1e9:.1f} TFLOP/s-equivalent")

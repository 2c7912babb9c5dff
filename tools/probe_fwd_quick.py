"""Quick forward timing + sanity for one build of the library (select with FA_B200_LIB=<path>): C2, causal, C3-fwd, short
causal, hd64, fp8 C4; max|out - FA2| on the C2 / causal shapes as a smoke check.  One line per config."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa  # noqa: E402

tag = os.environ.get("FA_TAG", os.path.basename(os.environ.get("FA_B200_LIB", "default")))


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


def rnd(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device="cuda", dtype=torch.bfloat16).to(dtype)


torch.manual_seed(0)
try:
    import importlib
    real = importlib.import_module("flash_attn_2_cuda")
except Exception:  # noqa: BLE001
    real = None
res = {}
for name, b, h, hk, s, d, causal in (("C2", 4, 16, 16, 8192, 128, False), ("causal8k", 4, 16, 16, 8192, 128, True),
                                      ("C3fwd", 2, 32, 4, 16384, 128, True), ("causal1k", 32, 16, 16, 1024, 128, True),
                                      ("causal2k", 16, 16, 16, 2048, 128, True), ("hd64", 4, 32, 32, 8192, 64, False)):
    q, k, v = rnd(b, s, h, d), rnd(b, s, hk, d), rnd(b, s, hk, d)
    ms = bench(lambda: fa.flash_attn_func(q, k, v, causal=causal))
    fl = 4 * b * h * s * s * d / (2 if causal else 1)
    err = None
    if real is not None and name in ("C2", "causal8k"):
        o = fa.flash_attn_func(q, k, v, causal=causal)
        r = real.fwd(q, k, v, None, None, 0.0, d ** -0.5, causal, -1, -1, 0.0, False, None)[0]
        err = (o.float() - r.float()).abs().max().item()
    res[name] = round(fl / ms / 1e9, 1)
    print(f"[{tag}] {name}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s" + (f" max|o-FA2| {err:.2e}" if err is not None else ""), flush=True)
f8 = torch.float8_e4m3fn
q, k, v = (rnd(1, 32768, 16, 128, dtype=f8) for _ in range(3))
ms = bench(lambda: fa.flash_attn_func(q, k, v, causal=True, window_size=(4096, 0)))
pairs = sum(min(i, 4096) + 1 for i in range(32768))
res["C4fp8"] = round(4 * 16 * pairs * 128 / ms / 1e9, 1)
print(f"[{tag}] C4 fp8: {ms:.3f} ms {res['C4fp8']} TFLOP/s", flush=True)
print(json.dumps({"tag": tag, **res}), flush=True)

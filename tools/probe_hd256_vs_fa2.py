"""Head dim 256 (and 192) forward / backward: this library vs the installed FA2 2.8.3 wheel on the same B200
(the reference's benchmark shape: 16k tokens, hidden 2048 -> 8 heads of 256, benchmarks/benchmark_flash_attention.py:72-76)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.baseline_probe import bench  # noqa: E402


def run(tag, fn):
    for d, h in ((256, 8), (192, 8)):
        for causal in (False, True):
            b, s = 2, 8192
            torch.manual_seed(0)
            q, k, v = (torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
            do = torch.randn(b, s, h, d, device="cuda", dtype=torch.bfloat16)
            fl = 4 * b * h * s * s * d / (2 if causal else 1)
            try:
                ms = bench(lambda: fn(q, k, v, causal=causal), n=5, warm=2)
                out = fn(q, k, v, causal=causal)
                ms_b = bench(lambda: out.backward(do, retain_graph=True), n=5, warm=2)
                print(f"[{tag}] d{d} causal={int(causal)} fwd {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s | bwd {ms_b:.3f} ms "
                      f"{2.5 * fl / ms_b / 1e9:.0f} TFLOP/s", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"[{tag}] d{d} causal={int(causal)} FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)


if __name__ == "__main__":
    import flash_attention_b200 as fab
    run("fa_b200", fab.flash_attn_func)
    if len(sys.argv) > 1 and sys.argv[1] == "ours":
        sys.exit(0)
    try:
        from flash_attn import flash_attn_func
        run("FA2-2.8.3", flash_attn_func)
    except Exception as e:  # noqa: BLE001
        print("FA2 unavailable:", e)

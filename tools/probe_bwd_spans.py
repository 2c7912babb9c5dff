"""Where does the pair backward lose time between CTAs?  (needs the -DFA_BWD2_TS=1 build)
Records (smid, entry, exit) of every CTA of one backward launch and reports busy / idle time per SM."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from flash_attention_b200 import _lib
from flash_attention_b200 import flash_attn_2_cuda as op
from tests.util import make_qkv

causal = len(sys.argv) > 1 and sys.argv[1] == "causal"
b, s, h, d = 4, 8192, 16, 128
q, k, v = make_qkv(b, s, s, h, h, d)
do = torch.randn_like(q)
scale = d ** -0.5
out, lse, _, _ = op.fwd(q, k, v, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
lib = _lib.load()
lib.fa_b200_debug_read_spans.argtypes = [C.c_void_p, C.c_int, C.c_int]
for _ in range(2):
    op.bwd(do, q, k, v, out, lse, dq, dk, dv, None, 0.0, scale, causal, -1, -1, 0.0, False, None, None)
torch.cuda.synchronize()
lib.fa_b200_debug_read_spans(None, 0, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
op.bwd(do, q, k, v, out, lse, dq, dk, dv, None, 0.0, scale, causal, -1, -1, 0.0, False, None, None)
e1.record()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (3 * 8192))()
n = lib.fa_b200_debug_read_spans(buf, 8192, 0)
rec = torch.tensor(list(buf[: 3 * n]), dtype=torch.int64).view(n, 3)
t0, t1 = int(rec[:, 1].min()), int(rec[:, 2].max())
print(f"causal={causal}: {n} CTAs, kernel span {(t1 - t0) / 1e3:.1f} us (whole bwd call {e0.elapsed_time(e1) * 1e3:.1f} us)")
by_sm = {}
for smid, a, z in rec.tolist():
    by_sm.setdefault(smid, []).append((a, z))
busy, gaps, first, last = [], [], [], []
for smid, iv in by_sm.items():
    iv.sort()
    busy.append(sum(z - a for a, z in iv))
    gaps += [iv[i + 1][0] - iv[i][1] for i in range(len(iv) - 1)]
    first.append(iv[0][0] - t0)
    last.append(t1 - iv[-1][1])
g = torch.tensor(gaps, dtype=torch.float64)
print(f"SMs {len(by_sm)}; CTAs per SM {n / len(by_sm):.1f}; mean busy {sum(busy) / len(busy) / 1e3:.1f} us = {100 * sum(busy) / len(busy) / (t1 - t0):.1f} % of the span")
print(f"gap between consecutive CTAs on an SM: mean {g.mean() / 1e3:.2f} us, median {g.median() / 1e3:.2f} us, p90 {g.quantile(0.9) / 1e3:.2f} us, total per SM {g.sum() / len(by_sm) / 1e3:.1f} us")
print(f"start skew (first CTA entry after kernel start): mean {sum(first) / len(first) / 1e3:.2f} us; tail idle (kernel end - last CTA exit): mean {sum(last) / len(last) / 1e3:.1f} us, max {max(last) / 1e3:.1f} us")
dur = torch.tensor([z - a for iv in by_sm.values() for a, z in iv], dtype=torch.float64)
print(f"CTA duration: mean {dur.mean() / 1e3:.1f} us, min {dur.min() / 1e3:.1f}, max {dur.max() / 1e3:.1f}")

"""Per-CTA work-item timeline of the persistent forward (needs the -DFA_FWD_TS=1 build, FA_B200_LIB=.../lib_ts.so):
for each CTA the globaltimer at entry / exit of every item of its static list -> per-SM item durations, the gaps between
items, and how far apart the CTAs finish (what a dynamic scheduler could recover)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from flash_attention_b200 import _lib
from tests.util import make_qkv

causal = len(sys.argv) > 1 and sys.argv[1] == "causal"
q, k, v = make_qkv(4, 8192, 8192, 16, 16, 128)
for _ in range(3):
    o = fab.flash_attn_func(q, k, v, causal=causal)
torch.cuda.synchronize()
lib = _lib.load()
ts = (C.c_ulonglong * (192 * 64 * 2))()
sm = (C.c_uint * 192)()
assert lib.fa_b200_debug_read_fwd_items(ts, sm) == 0
t = torch.tensor(list(ts), dtype=torch.int64).view(192, 64, 2)
G = 148
t = t[:G]
used = t[:, :, 1] > 0
t0 = int(t[:, 0, 0].min())
dur = (t[:, :, 1] - t[:, :, 0]).float() / 1e3
n_items = used.sum(1)
end = (torch.stack([t[c, int(n_items[c]) - 1, 1] for c in range(G)]) - t0).float()   # (subtract in int64: ns since boot)
start = (t[:, 0, 0] - t0).float()
print(f"causal={causal}: items per CTA {int(n_items.min())}..{int(n_items.max())}; kernel span {float(end.max()) / 1e3:.1f} us")
print(f"first-item entry: min {float(start.min()) / 1e3:.2f} us max {float(start.max()) / 1e3:.2f} us")
print(f"CTA finish time: min {float(end.min()) / 1e3:.1f} us  median {float(end.median()) / 1e3:.1f}  max {float(end.max()) / 1e3:.1f}")
d = dur[used]
print(f"item duration us: mean {float(d.mean()):.2f} min {float(d.min()):.2f} max {float(d.max()):.2f} std {float(d.std()):.2f}")
gaps = []
for c in range(G):
    for r in range(1, int(n_items[c])):
        gaps.append(float(t[c, r, 0] - t[c, r - 1, 1]) / 1e3)
g = torch.tensor(gaps)
print(f"gap between items us: mean {float(g.mean()):.2f} max {float(g.max()):.2f}")
per_cta = torch.stack([dur[c, : int(n_items[c])].mean() for c in range(G)])
order = per_cta.argsort()
print("fastest CTAs (sm, mean item us):", [(int(sm[int(c)]), round(float(per_cta[c]), 2)) for c in order[:8]])
print("slowest CTAs (sm, mean item us):", [(int(sm[int(c)]), round(float(per_cta[c]), 2)) for c in order[-8:]])
busy = torch.stack([dur[c, : int(n_items[c])].sum() for c in range(G)])
print(f"sum of item durations per CTA us: min {float(busy.min()):.1f} max {float(busy.max()):.1f}; ideal balanced span {float(busy.sum()) / G:.1f} us")
if not causal:
    by_round = [float(dur[:, r][used[:, r]].mean()) for r in range(int(n_items.max()))]
    print("mean item duration by round:", [round(x, 1) for x in by_round])

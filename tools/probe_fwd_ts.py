"""Pipeline timeline of the forward kernel (needs the -DFA_FWD_TS=1 build)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from flash_attention_b200 import _lib
from tests.util import make_qkv

q, k, v = make_qkv(4, 8192, 8192, 16, 16, 128)
for _ in range(3):
    o = fab.flash_attn_func(q, k, v)
torch.cuda.synchronize()
lib = _lib.load()
n = 16 * 16
buf = (C.c_ulonglong * n)()
lib.fa_b200_debug_read_fwd_ts.argtypes = [C.c_void_p, C.c_int]
assert lib.fa_b200_debug_read_fwd_ts(buf, n) == 0
ts = torch.tensor(list(buf), dtype=torch.int64).view(16, 16)
names = ["m:P0 half ok", "m:P0 full ok", "m:PV0 issued", "m:S0 issued", "m:P1 half ok", "m:P1 full ok", "m:PV1 issued", "m:S1 issued",
         "s0:S ok", "s0:max done", "s0:P half", "s0:P full", "s1:S ok", "s1:max done", "s1:P half", "s1:P full"]
for it in range(2, 10):
    base = int(ts[it, 0])
    prev = int(ts[it - 1, 0])
    print(f"it {it + 8}: period {base - prev} | " + " ".join(f"{nm}={int(ts[it, e]) - base:+d}" for e, nm in enumerate(names) if int(ts[it, e])))

"""Pipeline timeline of the wide-S forward kernel (needs a -DFA_FWD_TS=1 build; FA_B200_FWD_WIDE=1)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FA_B200_FWD_WIDE"] = "1"
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
names = ["m:QK issued", "m:P seen", "m:PV issued", "-",
         "a:S seen", "a:ld half", "a:max", "a:exch", "a:exp done", "a:P arrive",
         "b:S seen", "b:ld half", "b:max", "b:exch", "b:exp done", "b:P arrive"]
for it in range(2, 10):
    base = int(ts[it, 4])
    prev = int(ts[it - 1, 4])
    print(f"blk {it + 8}: period {base - prev} | " + " ".join(f"{nm}={int(ts[it, e]) - base:+d}" for e, nm in enumerate(names) if int(ts[it, e])))

"""Pipeline timeline of the CTA-pair backward kernel (needs the -DFA_BWD2_TS=1 build, FA_B200_BWD_PAIR=1)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from flash_attention_b200 import _lib
from tests.util import make_qkv

causal = len(sys.argv) > 1 and sys.argv[1] == "causal"
q, k, v = make_qkv(1, 8192, 8192, 2, 2, 128)
do = torch.randn_like(q)
q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
for _ in range(2):
    o = fab.flash_attn_func(q, k, v, causal=causal)
    o.backward(do)
torch.cuda.synchronize()
lib = _lib.load()
n = 2 * 16 * 32
buf = (C.c_ulonglong * n)()
lib.fa_b200_debug_read_ts.argtypes = [C.c_void_p, C.c_int]
assert lib.fa_b200_debug_read_ts(buf, n) == 0
ts = torch.tensor(list(buf), dtype=torch.int64).view(2, 16, 32)
names = ["P_F ok", "dV issued", "F_QB ok", "DS_F ok", "DS_IN ok", "DS_PEER ok", "DQ_E ok", "dP issued", "c:S_F ok", "c:P arrive",
         "c:DP_F ok", "c:dS arrive", "d:DQ_F ok", "d:DQ_E arrive", "c:dS math done", "c:dS stored", "c:after bar", "c:stores issued", "w4:math done", "w4:after bar", "w4:stored", "MMA dV done", "MMA S(n+1) done", "MMA dK done", "MMA dQ done", "MMA dP(n+1) done", "c:tmem st issued"]
for rank in range(2):
    print(f"=== rank {rank} (clock64 relative to the leader's 'P_F ok' of the same iteration; same-SM events only are comparable)")
    for it in range(2, 10):
        base = int(ts[rank, it, 0]) if rank == 0 else int(ts[rank, it, 8])
        row = []
        for e, nm in enumerate(names):
            if e in (1, 2, 4, 5, 18, 19, 20) and len(sys.argv) > 2:
                continue
            t = int(ts[rank, it, e])
            if t:
                row.append(f"{nm}={t - base:+d}")
        prev = int(ts[rank, it - 1, 0]) if rank == 0 else int(ts[rank, it - 1, 8])
        print(f"it {it + 8}: period {base - prev} | " + " ".join(row))

x = ts[0, 15]
names_x = ["entry", "before cluster sync", "after cluster sync", "K/V landed", "Q landed", "last MMA issued", "all MMA done (compute sees)",
           "epilogue stores issued", "after final cluster sync"]
print("once-per-CTA stamps (leader, clk since entry):", ", ".join(f"{n}={int(x[i]) - int(x[0])}" for i, n in enumerate(names_x) if int(x[i])))

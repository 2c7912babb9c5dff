import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from tests.util import make_qkv, fwd_tolerances
for kval, col in ((300.0, 13), (450.0, 200)):
    q, k, v = make_qkv(1, 1536, 1536, 2, 2, 128)
    q = q.clone(); k = k.clone(); q[..., 0] = 4.0; k[:, col, :, 0] = kval
    ref32, lse_ref, bound = fwd_tolerances(q, k, v)
    for wide in ("1", "0"):
        os.environ["FA_B200_FWD_WIDE"] = wide
        o, l, _ = fab.flash_attn_func(q, k, v, return_attn_probs=True)
        e = (o.float().cpu() - ref32).abs()
        idx = (e == e.max()).nonzero()[0].tolist()
        print(f"kval {kval} col {col} wide={wide}: err {e.max().item():.3e} bound {bound:.3e} at {idx} out {o[tuple(idx)].item()} ref {ref32[tuple(idx)].item()} v {v[0, col, idx[2], idx[3]].item()}  lse err {(l.float().cpu() - lse_ref).abs().max().item():.3e}  rows>bound {(e.amax(-1) > bound).sum().item()}")

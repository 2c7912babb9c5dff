"""Randomised stress of the backward (pair kernel by default): many small random shapes / masks, each run 3 times.
Checks (a) the 3x oracle rule, (b) dK/dV bitwise repeatable always, dQ bitwise repeatable with deterministic=True.
usage: stress_bwd.py [seed] [n_configs] [head dims, comma separated]   e.g. `stress_bwd.py 5 120 160,192,224,256`
for the head-dim 129..256 kernel (the default list keeps the sequence that tests/test_bwd_stress_gpu.py pins)."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from tests.util import bwd_tolerances, make_qkv

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 60
head_dims = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [128, 128, 128, 96, 64]
bad = 0
for ci in range(n_cfg):
    b = rng.choice([1, 1, 2, 3])
    hk = rng.choice([1, 2, 3])
    g = rng.choice([1, 1, 2, 4])
    h = hk * g
    d = rng.choice(head_dims)
    sq = rng.choice([rng.randint(1, 300), rng.randint(300, 1400), 128 * rng.randint(1, 9), 256 * rng.randint(1, 5)])
    sk = rng.choice([sq, rng.randint(1, 1400), 128 * rng.randint(1, 10), 256 * rng.randint(1, 5) + rng.choice([0, 1, 127])])
    mode = rng.choice(["none", "causal", "local", "causal_local"])
    causal = mode in ("causal", "causal_local")
    window = (-1, -1)
    # windows stay below seqlen_k like in the reference's tests (tests/test_flash_attn.py:945): at >= seqlen_k the reference's
    # host code turns a window limit into "infinite" (flash_api.cpp:414-419), which its python attention_ref does not do
    wmax = max(0, min(300, sk - 1))
    if mode == "local":
        window = (rng.randint(0, wmax), rng.randint(0, wmax))
    if mode == "causal_local":
        window = (rng.randint(0, wmax), -1)
    det = rng.random() < 0.5
    dtype = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype, seed=ci)
    do = torch.randn(b, sq, h, d, device="cuda", dtype=dtype)
    res = []
    for rep in range(3):
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = fab.flash_attn_func(qq, kk, vv, causal=causal, window_size=window, deterministic=det)
        out.backward(do)
        res.append((qq.grad, kk.grad, vv.grad))
    torch.cuda.synchronize()
    msgs = []
    tol = bwd_tolerances(q, k, v, do, causal=causal, window=window)
    for name, gt in zip(("dq", "dk", "dv"), res[0]):
        ref, bound = tol[name]
        err = (gt.float().cpu() - ref).abs().max().item()
        if not (err <= bound) or torch.isnan(gt).any():
            msgs.append(f"{name} err {err:.3e} > {bound:.3e}")
    for rep in (1, 2):
        if not torch.equal(res[0][1], res[rep][1]):
            msgs.append("dk not repeatable")
        if not torch.equal(res[0][2], res[rep][2]):
            msgs.append("dv not repeatable")
        if det and not torch.equal(res[0][0], res[rep][0]):
            msgs.append("dq not repeatable (deterministic)")
    tag = f"cfg {ci}: b{b} sq{sq} sk{sk} h{h}/{hk} d{d} {mode} w{window} det={det} {str(dtype)[6:]}"
    if msgs:
        bad += 1
        print("FAIL", tag, "|", "; ".join(sorted(set(msgs))), flush=True)
    elif ci % 10 == 0:
        print("ok  ", tag, flush=True)
print(f"stress done: {n_cfg - bad}/{n_cfg} configurations clean")

"""Randomised stress of the forward: shapes (incl. long sequences that take the chunked block order), GQA, masks, ALiBi,
softcap, dtypes; each checked with the reference's 2x rule against the oracle (out and LSE)."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from tests.util import check_fwd, make_qkv

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bad = 0
for ci in range(n_cfg):
    long_ = rng.random() < 0.25
    b = 1 if long_ else rng.choice([1, 2, 3])
    hk = 1 if long_ else rng.choice([1, 2, 3])
    g = rng.choice([1, 1, 2, 4])
    h = hk * g
    d = rng.choice([128, 128, 64, 96, 256]) if not long_ else rng.choice([64, 128])
    if long_:
        sq = rng.randint(2100, 5200)
        sk = rng.choice([sq, sq + rng.randint(0, 300)])
    else:
        sq = rng.choice([rng.randint(1, 300), rng.randint(300, 1400), 128 * rng.randint(1, 9)])
        sk = rng.choice([sq, rng.randint(1, 1400), 128 * rng.randint(1, 10)])
    mode = rng.choice(["none", "causal", "local", "causal_local"])
    causal = mode in ("causal", "causal_local")
    window = (-1, -1)
    wmax = max(0, min(400, sk - 1))
    if mode == "local":
        window = (rng.randint(0, wmax), rng.randint(0, wmax))
    if mode == "causal_local":
        window = (rng.randint(0, wmax), -1)
    softcap = rng.choice([0.0, 0.0, 0.0, 20.0])
    alibi = None
    if rng.random() < 0.2:
        alibi = (torch.rand(b, h, device="cuda") * 0.3).float()
    dtype = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype, seed=1000 + ci)
    out, lse, _ = fab.flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=alibi,
                                      return_attn_probs=True)
    torch.cuda.synchronize()
    tag = f"cfg {ci}: b{b} sq{sq} sk{sk} h{h}/{hk} d{d} {mode} w{window} softcap={softcap} alibi={alibi is not None} {str(dtype)[6:]}"
    try:
        check_fwd(out, lse, q, k, v, name=tag, causal=causal, window=window, softcap=softcap, alibi_slopes=alibi)
        if ci % 10 == 0:
            print("ok  ", tag, flush=True)
    except AssertionError as e:
        bad += 1
        print("FAIL", str(e)[:300], flush=True)
print(f"stress done: {n_cfg - bad}/{n_cfg} configurations clean")

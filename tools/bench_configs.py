"""Measure the BASELINE.json configs C2-C5 on one B200 (CUDA events, 3 warm-up + 10 timed, inputs > L2 or
HBM-streaming).  Prints one JSON line per config.  usage: python tools/bench_configs.py [ours|fa2]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "ours"
if which == "ours":
    import flash_attention_b200 as fa
else:
    import flash_attn as fa


def timeit(f, n=10, warm=3):
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


def rnd(*shape):
    return torch.randn(*shape, device="cuda", dtype=torch.bfloat16)


res = []
# C2: bf16 fwd non-causal b4 h16 s8192 d128
q, k, v = rnd(4, 8192, 16, 128), rnd(4, 8192, 16, 128), rnd(4, 8192, 16, 128)
ms = timeit(lambda: fa.flash_attn_func(q, k, v))
fl = 4 * 4 * 16 * 8192 * 8192 * 128
res.append(dict(cfg="C2 bf16 fwd non-causal b4 h16 s8192 d128", ms=ms, tflops=fl / ms / 1e9))
# headline: fwd+bwd same shape, causal and not
for causal in (False, True):
    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
    do = torch.randn_like(q)
    def fb():
        o = fa.flash_attn_func(qg, kg, vg, causal=causal)
        o.backward(do)
        qg.grad = kg.grad = vg.grad = None
    ms = timeit(fb)
    res.append(dict(cfg=f"headline bf16 fwd+bwd {'causal' if causal else 'non-causal'} b4 h16 s8192 d128", ms=ms,
                    tflops=3.5 * fl / (2 if causal else 1) / ms / 1e9))
del q, k, v, qg, kg, vg
# C3: bf16 fwd+bwd causal GQA 32:4 s16384 d128, b=2
q, k, v = rnd(2, 16384, 32, 128), rnd(2, 16384, 4, 128), rnd(2, 16384, 4, 128)
qg, kg, vg = (t.requires_grad_(True) for t in (q, k, v))
do = torch.randn_like(q)
def fb3():
    o = fa.flash_attn_func(qg, kg, vg, causal=True)
    o.backward(do)
    qg.grad = kg.grad = vg.grad = None
ms = timeit(fb3, n=5)
fl3 = 4 * 2 * 32 * 16384 * 16384 * 128 / 2
res.append(dict(cfg="C3 bf16 fwd+bwd causal GQA 32:4 s16384 d128 b2", ms=ms, tflops=3.5 * fl3 / ms / 1e9))
del q, k, v, qg, kg, vg, do
# C4: fp8 e4m3 fwd s32768 d128 window 4096 (b1 h16)
if which == "ours":
    f8 = torch.float8_e4m3fn
    q, k, v = rnd(1, 32768, 16, 128).to(f8), rnd(1, 32768, 16, 128).to(f8), rnd(1, 32768, 16, 128).to(f8)
    ms = timeit(lambda: fa.flash_attn_func(q, k, v, causal=True, window_size=(4096, 0)))
    pairs = sum(min(i, 4096) + 1 for i in range(32768))
    res.append(dict(cfg="C4 fp8 e4m3 fwd b1 h16 s32768 d128 window(4096,0)", ms=ms, tflops=4 * 16 * pairs * 128 / ms / 1e9))
    del q, k, v
# C5: decode b64 sq1 h32 hk4 d128 kv65536 paged(256) split-KV
b, h, hk, d, page, kv = 64, 32, 4, 128, 256, 65536
nb = kv // page
kc = rnd(b * nb, page, hk, d)
vc = rnd(b * nb, page, hk, d)
bt = torch.randperm(b * nb, device="cuda").int().view(b, nb)
cs = torch.full((b,), kv, dtype=torch.int32, device="cuda")
q = rnd(b, 1, h, d)
for ns in ((0,) if which != "ours" else (0, 1, 2, 4)):
    ms = timeit(lambda: fa.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cs, block_table=bt, num_splits=ns), n=5)
    byts = 2 * b * kv * hk * d * 2
    res.append(dict(cfg=f"C5 decode b64 sq1 h32/4 d128 kv65536 paged256 num_splits={ns}", ms=ms, gbps=byts / ms / 1e6,
                    frac_hbm_6572=byts / ms / 1e6 / 6572.2))
# ragged varlen (packed, causal): 32k tokens in sequences from 8192 down to 1 token -- exercises the persistent kernels'
# work-list scheduling (a max_seqlen grid would launch 24 x 64 m-blocks per head for 246 that exist)
lens = [8192, 4096, 4096, 2048, 2048, 2048, 1024, 1024, 1024, 1024, 512, 512, 512, 512, 300, 300, 211, 173, 128, 64, 64, 33, 17, 1]
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
tot = sum(lens)
qv, kv_, vv = (rnd(tot, 16, 128).requires_grad_(True) for _ in range(3))
dov = rnd(tot, 16, 128)
flv = sum(4 * 16 * (l * (l + 1) // 2) * 128 for l in lens)   # causal pairs per sequence


def fbv():
    o = fa.flash_attn_varlen_func(qv, kv_, vv, cu, cu, max(lens), max(lens), causal=True)
    o.backward(dov)
    qv.grad = kv_.grad = vv.grad = None


ms = timeit(lambda: fa.flash_attn_varlen_func(qv, kv_, vv, cu, cu, max(lens), max(lens), causal=True))
res.append(dict(cfg="ragged varlen causal fwd, 24 sequences 1..8192 tokens (32k total) h16 d128", ms=ms, tflops=flv / ms / 1e9))
ms = timeit(fbv, n=5)
res.append(dict(cfg="ragged varlen causal fwd+bwd, 24 sequences 1..8192 tokens (32k total) h16 d128", ms=ms, tflops=3.5 * flv / ms / 1e9))
for r in res:
    print(json.dumps(r), flush=True)

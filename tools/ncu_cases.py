"""One named kernel configuration, launched a few times -- the target of `ncu --set full -k regex:<kernel>` captures for
the kernels that had none in round 1: fp8 C4, decode C5 + combine, head-dim 256 backward, head dim 64
forward / backward, ragged varlen.   usage: python tools/ncu_cases.py <case> [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fa  # noqa: E402

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3


def rnd(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device="cuda", dtype=torch.bfloat16).to(dtype)


def fwd_bwd(b, s, h, hk, d, causal):
    q, k, v = (rnd(b, s, hh, d).requires_grad_(True) for hh in (h, hk, hk))
    do = rnd(b, s, h, d)

    def f():
        fa.flash_attn_func(q, k, v, causal=causal).backward(do)
        q.grad = k.grad = v.grad = None
    return f


if case == "fp8_c4":      # BASELINE C4: fp8 e4m3 fwd b1 h16 s32768 d128 window (4096, 0)
    f8 = torch.float8_e4m3fn
    q, k, v = (rnd(1, 32768, 16, 128, dtype=f8) for _ in range(3))
    f = lambda: fa.flash_attn_func(q, k, v, causal=True, window_size=(4096, 0))
elif case == "bf16_c4":   # same mask in bf16 (the comparison the fp8 kernel has to win)
    q, k, v = (rnd(1, 32768, 16, 128) for _ in range(3))
    f = lambda: fa.flash_attn_func(q, k, v, causal=True, window_size=(4096, 0))
elif case == "decode_c5":  # BASELINE C5 at b = 16 (2 GiB of cache: ncu replays every launch ~40 times)
    b, h, hk, d, page, kv = 16, 32, 4, 128, 256, 65536
    nb = kv // page
    kc, vc = rnd(b * nb, page, hk, d), rnd(b * nb, page, hk, d)
    bt = torch.randperm(b * nb, device="cuda").int().view(b, nb)
    cs = torch.full((b,), kv, dtype=torch.int32, device="cuda")
    q = rnd(b, 1, h, d)
    f = lambda: fa.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cs, block_table=bt)
elif case == "bwd_hd256":
    f = fwd_bwd(2, 8192, 8, 8, 256, False)
elif case == "hd64":
    f = fwd_bwd(4, 8192, 32, 32, 64, False)
elif case == "hd64_causal":
    f = fwd_bwd(4, 8192, 32, 32, 64, True)
elif case == "headline":
    f = fwd_bwd(4, 8192, 16, 16, 128, False)
elif case == "headline_fwd":   # forward only (FA_B200_FWD_WIDE=1 selects the wide-S kernel)
    q, k, v = (rnd(4, 8192, 16, 128) for _ in range(3))
    f = lambda: fa.flash_attn_func(q, k, v)
elif case == "headline_causal":
    f = fwd_bwd(4, 8192, 16, 16, 128, True)
elif case == "c3":
    f = fwd_bwd(2, 16384, 32, 4, 128, True)
elif case == "varlen_ragged":   # 32k tokens in sequences of very different lengths, causal
    lens = [8192, 4096, 4096, 2048, 2048, 2048, 1024, 1024, 1024, 1024, 512, 512, 512, 512, 300, 300, 211, 173, 128, 64, 64, 33, 17, 1]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    tot = sum(lens)
    q, k, v = (rnd(tot, 16, 128).requires_grad_(True) for _ in range(3))
    do = rnd(tot, 16, 128)

    def f():
        fa.flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True).backward(do)
        q.grad = k.grad = v.grad = None
else:
    raise SystemExit(f"unknown case {case}")

for _ in range(reps):
    f()
torch.cuda.synchronize()
print("ran", case, reps)

"""Which gradient differs between two deterministic runs, and by how much."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from tests.util import make_qkv

b, sq, sk, h, hk, d = 2, 640, 900, 4, 2, 128
q, k, v = make_qkv(b, sq, sk, h, hk, d)
do = torch.randn(b, sq, h, d, device="cuda", dtype=torch.bfloat16)


def grads(det):
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = fab.flash_attn_func(qq, kk, vv, deterministic=det)
    out.backward(do)
    return qq.grad, kk.grad, vv.grad


for det in (True, False):
    g0 = grads(det)
    worst = [0, 0, 0]
    nbad = [0, 0, 0]
    where = [None, None, None]
    for _ in range(10):
        g = grads(det)
        for i, (a, c) in enumerate(zip(g0, g)):
            diff = (a.float() - c.float()).abs()
            if diff.max().item() > 0:
                nbad[i] += 1
                if diff.max().item() > worst[i]:
                    worst[i] = diff.max().item()
                    idx = torch.nonzero(diff > 0)
                    where[i] = (idx[:, 0].unique().tolist(), idx[:, 1].min().item(), idx[:, 1].max().item(), idx[:, 2].unique().tolist(),
                                idx.shape[0])
    print(f"deterministic={det}: runs differing dq/dk/dv = {nbad}, max abs diff = {worst}")
    for i, n in enumerate(("dq", "dk", "dv")):
        if where[i]:
            print(f"   {n}: batches {where[i][0]} rows {where[i][1]}..{where[i][2]} heads {where[i][3]} count {where[i][4]}")

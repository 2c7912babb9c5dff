"""GPU bring-up probe for the backward kernels. usage: python tools/probe_bwd.py <stage>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from oracle.attention_oracle import attention_oracle_bwd, attention_oracle_varlen
from tests.util import bwd_tolerances, make_qkv


def run(name, b, sq, sk, h, hk, d, dtype=torch.bfloat16, diag=False, **kw):
    q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype)
    g = torch.Generator(device="cpu").manual_seed(7)
    do = torch.randn(b, sq, h, d, generator=g).to(dtype).cuda()
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    kw_api = dict(causal=kw.get("causal", False), window_size=kw.get("window", (-1, -1)), softcap=kw.get("softcap", 0.0),
                  alibi_slopes=kw.get("alibi_slopes"))
    out = fab.flash_attn_func(q, k, v, **kw_api)
    out.backward(do)
    torch.cuda.synchronize()
    tol = bwd_tolerances(q.detach(), k.detach(), v.detach(), do, **kw)
    ok = True
    msg = []
    for n, g_ in (("dq", q.grad), ("dk", k.grad), ("dv", v.grad)):
        ref, bound = tol[n]
        err = (g_.float().cpu() - ref).abs().max().item()
        nan = torch.isnan(g_).any().item()
        good = (err <= bound) and not nan
        ok &= good
        msg.append(f"{n} err={err:.3e} bound={bound:.3e}{' NAN' if nan else ''}{'' if good else ' <<<'}")
        if diag and not good:
            e = (g_.float().cpu() - ref).abs()
            idx = torch.nonzero(e > bound)[:5].tolist()
            print(f"   {n} bad idx sample {idx}; got {[g_[tuple(i)].item() for i in idx]} ref {[ref[tuple(i)].item() for i in idx]}")
            # per-row error profile
            rows = e.amax(dim=(0, 2, 3))
            print(f"   {n} row err profile (first 16 of {rows.numel()}):", [round(x, 3) for x in rows[:16].tolist()], " rel ratio sum:",
                  (g_.float().cpu().abs().sum() / ref.abs().sum()).item())
    print(f"[{name}] {'OK  ' if ok else 'FAIL'} " + " | ".join(msg), flush=True)


def stage_small():
    run("1blk", 1, 128, 128, 1, 1, 128, diag=True)
    run("2x2", 1, 256, 256, 1, 1, 128, diag=True)
    run("multi", 2, 384, 640, 2, 2, 128, diag=True)
    run("ragged", 2, 200, 333, 3, 3, 128, diag=True)
    run("fp16", 1, 256, 512, 2, 2, 128, torch.float16)
    run("d64", 1, 256, 384, 2, 2, 64, diag=True)
    run("d96", 1, 130, 257, 2, 2, 96)
    run("d32", 1, 130, 257, 2, 2, 32)


def stage_masks():
    run("causal", 1, 512, 512, 2, 2, 128, causal=True, diag=True)
    run("causal sq<sk", 2, 200, 461, 2, 2, 128, causal=True)
    run("causal sq>sk", 1, 300, 140, 2, 2, 128, causal=True)
    run("local", 1, 400, 500, 2, 2, 128, window=(100, 37))
    run("local left only", 1, 400, 500, 2, 2, 128, window=(64, -1))
    run("causal+left", 1, 700, 700, 1, 1, 128, causal=True, window=(129, -1))
    run("softcap", 1, 256, 256, 2, 2, 128, softcap=15.0)
    sl = (torch.rand(2, 4, device="cuda") * 0.3).float()
    run("alibi", 2, 200, 333, 4, 4, 128, alibi_slopes=sl)
    run("alibi causal", 2, 200, 333, 4, 4, 128, alibi_slopes=sl, causal=True)
    run("alibi softcap", 2, 200, 333, 4, 4, 128, alibi_slopes=sl, softcap=20.0)


def stage_gqa():
    run("gqa 8:2", 2, 300, 300, 8, 2, 128, causal=True, diag=True)
    run("mqa", 1, 256, 512, 4, 1, 128)
    run("gqa d64", 1, 300, 200, 6, 2, 64, causal=True)


def stage_varlen():
    torch.manual_seed(0)
    h, hk, d = 4, 2, 128
    lens_q = [37, 256, 1, 300]
    lens_k = [100, 256, 513, 300]
    cu_q = torch.tensor([0] + list(torch.tensor(lens_q).cumsum(0)), dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0] + list(torch.tensor(lens_k).cumsum(0)), dtype=torch.int32, device="cuda")
    for causal in (False, True):
        q = torch.randn(sum(lens_q), h, d, device="cuda").bfloat16().requires_grad_(True)
        k = torch.randn(sum(lens_k), hk, d, device="cuda").bfloat16().requires_grad_(True)
        v = torch.randn(sum(lens_k), hk, d, device="cuda").bfloat16().requires_grad_(True)
        do = torch.randn(sum(lens_q), h, d, device="cuda").bfloat16()
        o = fab.flash_attn_varlen_func(q, k, v, cu_q, cu_k, max(lens_q), max(lens_k), causal=causal)
        o.backward(do)
        torch.cuda.synchronize()
        for i in range(len(lens_q)):
            a, b_ = int(cu_q[i]), int(cu_q[i + 1]); c, e = int(cu_k[i]), int(cu_k[i + 1])
            _, _, dq, dk, dv = attention_oracle_bwd(q[a:b_].detach().cpu().unsqueeze(0), k[c:e].detach().cpu().unsqueeze(0),
                                                    v[c:e].detach().cpu().unsqueeze(0), do[a:b_].cpu().unsqueeze(0), causal=causal)
            print(f"[varlen causal={causal} seq{i}] dq {(q.grad[a:b_].float().cpu() - dq[0]).abs().max().item():.3e} "
                  f"dk {(k.grad[c:e].float().cpu() - dk[0]).abs().max().item():.3e} dv {(v.grad[c:e].float().cpu() - dv[0]).abs().max().item():.3e} "
                  f"(ref max {dq.abs().max().item():.2f} {dk.abs().max().item():.2f} {dv.abs().max().item():.2f})", flush=True)


def stage_time():
    for (b, h, hk, s, causal) in ((4, 16, 16, 8192, False), (4, 16, 16, 8192, True), (2, 32, 4, 16384, True)):
        q, k, v = make_qkv(b, s, s, h, hk, 128)
        do = torch.randn_like(q)
        q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
        def fwd():
            return fab.flash_attn_func(q, k, v, causal=causal)
        def fb():
            o = fwd(); o.backward(do); q.grad = k.grad = v.grad = None
        res = {}
        for nm, f, mult in (("fwd", fwd, 1.0), ("fwd+bwd", fb, 3.5)):
            for _ in range(3): f()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            n = 10
            for _ in range(n): f()
            ev[1].record(); torch.cuda.synchronize()
            res[nm] = ev[0].elapsed_time(ev[1]) / n
        fl = 4 * b * h * s * s * 128 / (2 if causal else 1)
        bw = res["fwd+bwd"] - res["fwd"]
        print(f"[time] b{b} h{h}/{hk} s{s} causal={causal}: fwd {res['fwd']:.3f} ms {fl / res['fwd'] / 1e9:.1f} TF | fwd+bwd {res['fwd+bwd']:.3f} ms "
              f"{3.5 * fl / res['fwd+bwd'] / 1e9:.1f} TF | bwd {bw:.3f} ms {2.5 * fl / bw / 1e9:.1f} TF", flush=True)


if __name__ == "__main__":
    st = sys.argv[1]
    print("==== stage", st, flush=True)
    globals()["stage_" + st]()
    torch.cuda.synchronize()
    print("==== stage", st, "done", flush=True)

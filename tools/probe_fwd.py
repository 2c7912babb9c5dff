"""GPU bring-up probe for the forward kernel: per-stage diagnostics (run under gpurun).
usage: python tools/probe_fwd.py <stage>    stages: s_tile, small, masks, gqa, varlen, split, paged, d64, time
"""
import ctypes as C
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from flash_attention_b200 import _lib, flash_attn_2_cuda as op
from oracle.attention_oracle import attention_oracle, attention_oracle_varlen, gather_paged_kv
from tests.util import make_qkv, check_fwd


def report(name, out, ref):
    err = (out.float().cpu() - ref.float().cpu()).abs()
    print(f"[{name}] max_err={err.max().item():.4e} mean_err={err.mean().item():.4e} ref_max={ref.abs().max().item():.3f} nan={torch.isnan(out).any().item()}", flush=True)
    return err.max().item()


def stage_s_tile():
    """raw S = Q K^T of CTA (0,0,0): validates TMA swizzle + K-major descriptors + TMEM ld layout"""
    lib = _lib.load()
    b, sq, sk, h, d = 1, 128, 128, 1, 128
    q, k, v = make_qkv(b, sq, sk, h, h, d)
    out = torch.empty_like(q)
    lse = torch.empty(b, h, sq, dtype=torch.float32, device="cuda")
    dbg = torch.full((128, 128), float("nan"), dtype=torch.float32, device="cuda")
    a = _lib.FwdArgs()
    a.q, a.k, a.v, a.out, a.softmax_lse = (C.c_void_p(t.data_ptr()) for t in (q, k, v, out, lse))
    for n, t in (("q", q), ("k", k), ("v", v), ("o", out)):
        setattr(a, n + "_batch_stride", t.stride(0)); setattr(a, n + "_row_stride", t.stride(1)); setattr(a, n + "_head_stride", t.stride(2))
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = b, sq, sk, h, h, d
    a.dtype = _lib.BF16
    a.softmax_scale = 1.0 / math.sqrt(d)
    a.window_size_left = a.window_size_right = -1
    a.num_splits = 1
    a.debug_s = C.c_void_p(dbg.data_ptr())
    rc = lib.fa_b200_mha_fwd(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    print("rc", rc, _lib.last_error() if rc else "")
    torch.cuda.synchronize()
    s_ref = q[0, :, 0].float() @ k[0, :, 0].float().T
    e = report("S tile", dbg, s_ref)
    if e > 1e-2:
        # which 16-wide K chunks made it into the sum?  least squares over per-chunk partial products
        parts = torch.stack([(q[0, :, 0, 16 * i:16 * i + 16].float() @ k[0, :, 0, 16 * i:16 * i + 16].float().T).flatten() for i in range(8)], 1)
        sol = torch.linalg.lstsq(parts.cpu(), dbg.flatten().cpu().nan_to_num().unsqueeze(1)).solution.flatten()
        print("  chunk coefficients:", [round(x, 3) for x in sol.tolist()])
        print("  S_dbg[0,:8]", dbg[0, :8].tolist()); print("  S_ref[0,:8]", s_ref[0, :8].tolist())
        print("  S_dbg[:8,0]", dbg[:8, 0].tolist()); print("  S_ref[:8,0]", s_ref[:8, 0].tolist())
        print("  vs S_ref^T:", (dbg - s_ref.T).abs().max().item())
    ref, lse_ref = attention_oracle(q.cpu(), k.cpu(), v.cpu())
    e2 = report("O (128x128)", out, ref)
    report("LSE", lse, lse_ref)
    if e2 > 3e-2:
        # P V diagnostics: try V transposed / chunk structure
        p = torch.softmax(s_ref * a.softmax_scale, -1).to(torch.bfloat16).float()
        parts = torch.stack([(p[:, 16 * i:16 * i + 16] @ v[0, 16 * i:16 * i + 16, 0].float()).flatten() for i in range(8)], 1)
        sol = torch.linalg.lstsq(parts.cpu(), out[0, :, 0].float().flatten().cpu().nan_to_num().unsqueeze(1)).solution.flatten()
        print("  PV chunk coefficients:", [round(x, 3) for x in sol.tolist()])
        print("  O[0,:8]", out[0, 0, 0, :8].tolist()); print("  ref[0,:8]", ref[0, 0, 0, :8].tolist())
        print("  O[0,64:72]", out[0, 0, 0, 64:72].tolist()); print("  ref[0,64:72]", ref[0, 0, 0, 64:72].tolist())


def run_dense(name, b, sq, sk, h, hk, d, dtype=torch.bfloat16, **kw):
    q, k, v = make_qkv(b, sq, sk, h, hk, d, dtype)
    kw_op = dict(kw)
    causal = kw_op.pop("causal", False)
    window = kw_op.pop("window", (-1, -1))
    softcap = kw_op.pop("softcap", 0.0)
    slopes = kw_op.pop("alibi_slopes", None)
    out = fab.flash_attn_func(q, k, v, causal=causal, window_size=window, softcap=softcap, alibi_slopes=slopes, return_attn_probs=True)
    o, lse = out[0], out[1]
    torch.cuda.synchronize()
    try:
        err, bound = check_fwd(o, lse, q, k, v, name=name, causal=causal, window=window, softcap=softcap, alibi_slopes=slopes)
        print(f"[{name}] OK err={err:.3e} bound={bound:.3e}", flush=True)
    except AssertionError as e:
        print(f"[{name}] FAIL {e}", flush=True)


def stage_small():
    run_dense("1tile", 1, 128, 128, 1, 1, 128)
    run_dense("2tiles", 1, 256, 256, 1, 1, 128)
    run_dense("multi-kv", 1, 256, 1024, 2, 2, 128)
    run_dense("ragged", 2, 200, 333, 3, 3, 128)
    run_dense("sq1", 2, 1, 300, 4, 4, 128)
    run_dense("fp16", 1, 256, 512, 2, 2, 128, torch.float16)
    run_dense("d64", 1, 256, 384, 2, 2, 64)
    run_dense("d96", 1, 130, 257, 2, 2, 96)
    run_dense("d32", 1, 130, 257, 2, 2, 32)


def stage_masks():
    run_dense("causal", 1, 512, 512, 2, 2, 128, causal=True)
    run_dense("causal sq<sk", 2, 200, 461, 2, 2, 128, causal=True)
    run_dense("causal sq>sk", 1, 300, 140, 2, 2, 128, causal=True)
    run_dense("local", 1, 400, 500, 2, 2, 128, window=(100, 37))
    run_dense("local left only", 1, 400, 500, 2, 2, 128, window=(64, -1))
    run_dense("causal+left", 1, 700, 700, 1, 1, 128, causal=True, window=(129, -1))
    run_dense("softcap", 1, 256, 256, 2, 2, 128, softcap=15.0)
    sl = (torch.rand(2, 4, device="cuda") * 0.3).float()
    run_dense("alibi", 2, 200, 333, 4, 4, 128, alibi_slopes=sl)
    run_dense("alibi causal", 2, 200, 333, 4, 4, 128, alibi_slopes=sl, causal=True)


def stage_gqa():
    run_dense("gqa 8:2", 2, 300, 300, 8, 2, 128, causal=True)
    run_dense("mqa decode packed", 3, 1, 777, 8, 1, 128)
    run_dense("gqa packed sq=5 causal", 2, 5, 260, 16, 2, 128, causal=True)
    run_dense("gqa packed sq=40", 1, 40, 300, 8, 2, 128, causal=True)


def stage_varlen():
    torch.manual_seed(0)
    h, hk, d = 4, 2, 128
    lens_q = [37, 256, 1, 300]
    lens_k = [100, 256, 513, 300]
    cu_q = torch.tensor([0] + list(torch.tensor(lens_q).cumsum(0)), dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0] + list(torch.tensor(lens_k).cumsum(0)), dtype=torch.int32, device="cuda")
    q = torch.randn(sum(lens_q), h, d, device="cuda").bfloat16()
    k = torch.randn(sum(lens_k), hk, d, device="cuda").bfloat16()
    v = torch.randn(sum(lens_k), hk, d, device="cuda").bfloat16()
    for causal in (False, True):
        o, lse, _ = fab.flash_attn_varlen_func(q, k, v, cu_q, cu_k, max(lens_q), max(lens_k), causal=causal, return_attn_probs=True)
        torch.cuda.synchronize()
        ref, lse_ref = attention_oracle_varlen(q.cpu(), k.cpu(), v.cpu(), cu_q.cpu(), cu_k.cpu(), max(lens_q), max(lens_k), causal=causal)
        report(f"varlen causal={causal}", o, ref)
        fin = torch.isfinite(lse_ref)
        print("   lse err", (lse.cpu()[fin] - lse_ref[fin]).abs().max().item())


def stage_split():
    q, k, v = make_qkv(2, 1, 4096, 8, 2, 128)
    for ns in (1, 2, 3, 7, 0):
        o, lse = fab.flash_attn_with_kvcache(q, k, v, num_splits=ns, return_softmax_lse=True)
        torch.cuda.synchronize()
        ref, lse_ref = attention_oracle(q.cpu(), k.cpu(), v.cpu())
        report(f"split ns={ns}", o, ref)
        print("   lse err", (lse.cpu() - lse_ref).abs().max().item())
    cs = torch.tensor([1000, 3333], dtype=torch.int32, device="cuda")
    for ns in (1, 4, 0):
        o, lse = fab.flash_attn_with_kvcache(q, k, v, cache_seqlens=cs, num_splits=ns, causal=True, return_softmax_lse=True)
        torch.cuda.synchronize()
        ref, lse_ref = attention_oracle(q.cpu(), k.cpu(), v.cpu(), k_lens=cs.cpu(), causal=True)
        report(f"cache_seqlens ns={ns}", o, ref)


def stage_paged():
    torch.manual_seed(0)
    b, h, hk, d, page = 3, 8, 2, 128, 256
    seqlens = torch.tensor([700, 256, 1500], dtype=torch.int32, device="cuda")
    nblk = 6
    num_pages = b * nblk
    kc = torch.randn(num_pages, page, hk, d, device="cuda").bfloat16()
    vc = torch.randn(num_pages, page, hk, d, device="cuda").bfloat16()
    bt = torch.randperm(num_pages, device="cuda").int().view(b, nblk)
    q = torch.randn(b, 1, h, d, device="cuda").bfloat16()
    kd = gather_paged_kv(kc.cpu(), bt.cpu(), nblk * page)
    vd = gather_paged_kv(vc.cpu(), bt.cpu(), nblk * page)
    for ns in (1, 0, 3):
        o = fab.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=seqlens, block_table=bt, num_splits=ns)
        torch.cuda.synchronize()
        ref, _ = attention_oracle(q.cpu(), kd, vd, k_lens=seqlens.cpu())
        report(f"paged ns={ns}", o, ref)


def stage_time():
    for (b, h, hk, s, causal) in ((4, 16, 16, 8192, False), (4, 16, 16, 8192, True), (2, 32, 4, 16384, True)):
        q, k, v = make_qkv(b, s, s, h, hk, 128)
        f = lambda: fab.flash_attn_func(q, k, v, causal=causal)
        for _ in range(3): f()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        n = 10
        for _ in range(n): f()
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        fl = 4 * b * h * s * s * 128 / (2 if causal else 1)
        print(f"[time] b{b} h{h}/{hk} s{s} causal={causal}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)




def stage_fp8():
    from oracle.attention_oracle import attention_oracle
    f8 = torch.float8_e4m3fn
    for name, b, sq, sk, h, hk, causal, window, desc in (
        ("fp8 1tile", 1, 128, 128, 1, 1, False, (-1, -1), False),
        ("fp8 multi", 2, 384, 640, 4, 2, False, (-1, -1), True),
        ("fp8 causal", 1, 512, 512, 2, 2, True, (-1, -1), True),
        ("fp8 window", 1, 1024, 1024, 2, 2, True, (256, 0), False),
        ("fp8 ragged", 2, 200, 333, 3, 3, False, (-1, -1), True),
    ):
        q, k, v = make_qkv(b, sq, sk, h, hk, 128)
        q8, k8, v8 = q.to(f8), k.to(f8), v.to(f8)
        qd = (torch.rand(b, hk, device="cuda") + 0.5).float() if desc else None
        kd = (torch.rand(b, hk, device="cuda") + 0.5).float() if desc else None
        vd = (torch.rand(b, hk, device="cuda") + 0.5).float() if desc else None
        out, lse, _ = fab.flash_attn_func(q8, k8, v8, causal=causal, window_size=window, return_attn_probs=True,
                                         q_descale=qd, k_descale=kd, v_descale=vd)
        torch.cuda.synchronize()
        c = lambda t: None if t is None else t.cpu()
        kw = dict(causal=causal, window=window, q_descale=c(qd), k_descale=c(kd), v_descale=c(vd))
        ref, lse_ref = attention_oracle(q8.cpu().float(), k8.cpu().float(), v8.cpu().float(), **kw)
        refq, _ = attention_oracle(q8.cpu().float(), k8.cpu().float(), v8.cpu().float(), p_dtype=f8, **kw)
        e = (out.float().cpu() - ref).abs().max().item()
        eq = (refq - ref).abs().max().item()
        fin = torch.isfinite(lse_ref)
        print(f"[{name}] err={e:.3e} (e4m3-P oracle err {eq:.3e}) ref_max={ref.abs().max().item():.2f} "
              f"lse_err={(lse.cpu()[fin] - lse_ref[fin]).abs().max().item():.2e} nan={torch.isnan(out).any().item()}", flush=True)
    # timing: BASELINE C4 = fp8 fwd b1 h16 s32768 d128 window (4096, 0)
    b, h, s = 1, 16, 32768
    q = torch.randn(b, s, h, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn_like(q); v = torch.randn_like(q)
    for nm, tq in (("fp8", (q.to(f8), k.to(f8), v.to(f8))), ("bf16", (q, k, v))):
        f = lambda: fab.flash_attn_func(*tq, causal=True, window_size=(4096, 0))
        for _ in range(3): f()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10): f()
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        pairs = sum(min(i, 4096) + 1 for i in range(s))
        print(f"[time C4 {nm}] {ms:.3f} ms {4 * b * h * pairs * 128 / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    st = sys.argv[1]
    print("==== stage", st, flush=True)
    globals()["stage_" + st]()
    torch.cuda.synchronize()
    print("==== stage", st, "done", flush=True)

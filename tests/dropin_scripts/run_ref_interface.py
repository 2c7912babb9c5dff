"""Runs in a FRESH interpreter (the reference interface files register torch custom ops under fixed names, so two versions
cannot live in one process): loads the REFERENCE's interface file given on the command line, unmodified, on top of this
package's compiled operator module and checks results against the CPU oracle.

  python run_ref_interface.py fa2 <path to flash_attn/flash_attn_interface.py>      (2.8.3 installed / 2.8.4 reference tree)
  python run_ref_interface.py fa3 <path to hopper/flash_attn_interface.py>
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import flash_attention_b200 as fab  # noqa: E402
from oracle.attention_oracle import attention_oracle, attention_oracle_bwd  # noqa: E402  (checker only)

kind, path = sys.argv[1], sys.argv[2]


def load(name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod   # torch.compile re-imports the module of a traced function by name
    spec.loader.exec_module(mod)
    return mod


def close(a, ref, tol, what):
    err = (a.float().cpu() - ref.float()).abs().max().item()
    assert err <= tol * max(1.0, ref.float().abs().max().item()), f"{what}: {err:.3e}"


torch.manual_seed(0)
dev = "cuda"
if kind == "fa2":
    ours = fab.install()
    iface = load("ref_flash_attn_interface")
    assert iface.flash_attn_gpu is ours and ours.__file__.endswith(".so")
    b, s, h, hk, d = 2, 384, 4, 2, 128
    q = torch.randn(b, s, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(b, s, hk, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(b, s, hk, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn_like(q)
    for causal in (False, True):
        out = iface.flash_attn_func(q, k, v, causal=causal)
        dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
        ro, _, rdq, rdk, rdv = attention_oracle_bwd(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(), causal=causal)
        close(out, ro, 2e-2, "out"); close(dq, rdq, 3e-2, "dq"); close(dk, rdk, 3e-2, "dk"); close(dv, rdv, 3e-2, "dv")
    # varlen (2.8.4 passes 22 positional arguments incl. num_splits; 2.8.3 passes 21)
    cu = torch.tensor([0, 100, 356, 387], dtype=torch.int32, device=dev)
    qv, kv, vv = (torch.randn(387, hh, d, device=dev, dtype=torch.bfloat16) for hh in (h, hk, hk))
    out = iface.flash_attn_varlen_func(qv, kv, vv, cu, cu, 256, 256, causal=True)
    for i in range(3):
        a, e = int(cu[i]), int(cu[i + 1])
        ro, _ = attention_oracle(qv[a:e].unsqueeze(0).cpu(), kv[a:e].unsqueeze(0).cpu(), vv[a:e].unsqueeze(0).cpu(), causal=True)
        close(out[a:e].unsqueeze(0), ro, 2e-2, f"varlen[{i}]")
    # kv-cache decode with an append
    kc = torch.randn(3, 512, hk, d, device=dev, dtype=torch.bfloat16)
    vc = torch.randn(3, 512, hk, d, device=dev, dtype=torch.bfloat16)
    cs = torch.tensor([100, 17, 300], dtype=torch.int32, device=dev)
    qd = torch.randn(3, 1, h, d, device=dev, dtype=torch.bfloat16)
    kn, vn = torch.randn(3, 1, hk, d, device=dev, dtype=torch.bfloat16), torch.randn(3, 1, hk, d, device=dev, dtype=torch.bfloat16)
    kc0, vc0 = kc.clone(), vc.clone()
    o = iface.flash_attn_with_kvcache(qd, kc, vc, k=kn, v=vn, cache_seqlens=cs, causal=True)
    for i in range(3):
        n = int(cs[i])
        assert torch.equal(kc[i, n], kn[i, 0]) and torch.equal(vc[i, n], vn[i, 0])   # appended in place
        kk = torch.cat([kc0[i, :n], kn[i]], 0).unsqueeze(0).cpu()
        vv_ = torch.cat([vc0[i, :n], vn[i]], 0).unsqueeze(0).cpu()
        ro, _ = attention_oracle(qd[i:i + 1].cpu(), kk, vv_, causal=True)
        close(o[i:i + 1], ro, 2e-2, f"kvcache[{i}]")
    # torch.compile(fullgraph=True) through the REFERENCE's custom ops + fakes on our kernels (a4)
    def fwd_only(q, k, v):
        return iface.flash_attn_func(q, k, v, causal=True)

    def step(f, q, k, v, do):
        out = f(q, k, v)
        return (out,) + torch.autograd.grad(out, (q, k, v), do)
    eager = step(fwd_only, q, k, v, do)
    comp = step(torch.compile(fwd_only, fullgraph=True), q, k, v, do)
    for a, c in zip(eager, comp):
        assert (a.float() - c.float()).abs().max().item() <= 2e-2
    print("OK fa2 interface", getattr(iface, "__file__", path))
else:
    fab.install_fa3()
    iface = load("ref_fa3_interface")
    b, s, h, hk, d = 2, 300, 4, 2, 128
    q = torch.randn(b, s, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(b, s, hk, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(b, s, hk, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn_like(q)
    res = iface.flash_attn_func(q, k, v, causal=True)
    out = res[0] if isinstance(res, tuple) else res
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    ro, _, rdq, rdk, rdv = attention_oracle_bwd(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(), causal=True)
    close(out, ro, 2e-2, "out"); close(dq, rdq, 3e-2, "dq"); close(dk, rdk, 3e-2, "dk"); close(dv, rdv, 3e-2, "dv")
    # fp8 with de-scales through the FA3 signature
    f8 = torch.float8_e4m3fn
    q8, k8, v8 = (t.detach().to(f8) for t in (q, k, v))
    ds = [(torch.rand(b, hk, device=dev) + 0.5).float() for _ in range(3)]
    res = iface.flash_attn_func(q8, k8, v8, causal=True, q_descale=ds[0], k_descale=ds[1], v_descale=ds[2])
    o8 = res[0] if isinstance(res, tuple) else res
    r8, _ = attention_oracle(q8.cpu().float(), k8.cpu().float(), v8.cpu().float(), causal=True, q_descale=ds[0].cpu(),
                             k_descale=ds[1].cpu(), v_descale=ds[2].cpu())
    close(o8, r8, 6e-2, "fp8 out")
    # kv cache: paged + append + seqused through page_table
    page, nb = 256, 4
    kc = torch.randn(2 * nb, page, hk, d, device=dev, dtype=torch.bfloat16)
    vc = torch.randn(2 * nb, page, hk, d, device=dev, dtype=torch.bfloat16)
    pt = torch.randperm(2 * nb, device=dev).int().view(2, nb)
    cs = torch.tensor([700, 33], dtype=torch.int32, device=dev)
    qd = torch.randn(2, 1, h, d, device=dev, dtype=torch.bfloat16)
    res = iface.flash_attn_with_kvcache(qd, kc, vc, cache_seqlens=cs, page_table=pt, causal=True)
    o = res[0] if isinstance(res, tuple) else res
    for i in range(2):
        n = int(cs[i])
        kk = kc[pt[i].long()].reshape(-1, hk, d)[:n].unsqueeze(0).cpu()
        vv_ = vc[pt[i].long()].reshape(-1, hk, d)[:n].unsqueeze(0).cpu()
        ro, _ = attention_oracle(qd[i:i + 1].cpu(), kk, vv_)
        close(o[i:i + 1], ro, 2e-2, f"fa3 kvcache[{i}]")
    # public combine op (hopper/flash_attn_interface.py flash_attn_combine)
    op = torch.randn(3, 2, 5, 4, 64, device=dev)
    lp = torch.randn(3, 2, 5, 4, device=dev)
    oc, lc = iface.flash_attn_combine(op, lp, out_dtype=torch.bfloat16)
    w = torch.softmax(lp, 0).unsqueeze(-1)
    assert (oc.float() - (w * op).sum(0)).abs().max().item() < 2e-2
    assert (lc - torch.logsumexp(lp, 0).transpose(1, 2)).abs().max().item() < 1e-4
    md = iface.get_scheduler_metadata(2, 1, nb * page, h, hk, d, cs, page_size=page)
    assert md.dtype == torch.int32 and md.numel() >= 3
    print("OK fa3 interface", path)

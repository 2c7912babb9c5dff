import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flash_attention_b200 as fab
from oracle.attention_oracle import attention_oracle
fa3 = fab.flash_attn_3
torch.manual_seed(0)
for (h, hk) in ((4, 2), (2, 2)):
    d = 128
    cu = torch.tensor([0, 300, 556], dtype=torch.int32, device="cuda")
    q = torch.randn(556, h, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(556, hk, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(556, hk, d, device="cuda", dtype=torch.bfloat16)
    for (uq, uk, causal) in (((300, 256), (300, 100), False), ((200, 256), (300, 256), False), ((200, 256), (300, 100), False),
                             ((200, 256), (300, 100), True), ((300, 256), (300, 100), True), ((200, 256), (300, 256), True)):
        tq = torch.tensor(uq, dtype=torch.int32, device="cuda"); tk = torch.tensor(uk, dtype=torch.int32, device="cuda")
        out = fa3.flash_attn_varlen_func(q, k, v, cu, cu, 300, 300, seqused_q=tq, seqused_k=tk, causal=causal)
        torch.cuda.synchronize()
        errs = []
        for i in range(2):
            a = int(cu[i]); nq, nk = uq[i], uk[i]
            ref, _ = attention_oracle(q[a:a+nq].cpu().unsqueeze(0), k[a:a+nk].cpu().unsqueeze(0), v[a:a+nk].cpu().unsqueeze(0), causal=causal)
            e = (out[a:a+nq].float().cpu() - ref[0].float()).abs()
            errs.append((round(e.max().item(), 4), int(e.amax(dim=(1, 2)).argmax())))
        print(f"h{h}/{hk} uq={uq} uk={uk} causal={causal}: (max err, worst row) per seq = {errs}", flush=True)

"""Host-buffer entry point: attention forward (+ backward) on tensors that live in pinned HOST memory.

The C ABI works on device pointers; a caller that owns host buffers (the reference-facing plugin call measured
as `e2e` by bench.py) would otherwise pay H2D -> compute -> D2H back to back.  This helper splits the batch into
chunks and runs a 3-stage pipeline on three CUDA streams (H2D of chunk i+1 | kernels of chunk i | D2H of chunk
i-1), so the step costs ~max(H2D, compute, D2H) instead of their sum.  Attention is independent per batch
element, so chunking does not change any result bit.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import flash_attn_2_cuda as _op


class HostAttentionPipeline:
    """Re-usable pipeline (streams, device staging buffers) for one problem shape."""

    def __init__(self, batch, seqlen_q, seqlen_k, nheads, nheads_k, headdim, dtype=torch.bfloat16, device="cuda",
                 chunks: Optional[int] = None, backward=True):
        self.dev = torch.device(device)
        self.b, self.bwd = batch, backward
        self.chunks = chunks or batch
        assert batch % self.chunks == 0
        cb = batch // self.chunks
        self.cb = cb
        mk = lambda s, h: [torch.empty(cb, s, h, headdim, dtype=dtype, device=self.dev) for _ in range(2)]
        # double-buffered device staging (chunk i+1 is uploaded while chunk i computes)
        self.q, self.k, self.v = mk(seqlen_q, nheads), mk(seqlen_k, nheads_k), mk(seqlen_k, nheads_k)
        self.do = mk(seqlen_q, nheads) if backward else None
        self.out = mk(seqlen_q, nheads)
        if backward:
            self.dq, self.dk, self.dv = mk(seqlen_q, nheads), mk(seqlen_k, nheads_k), mk(seqlen_k, nheads_k)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.scale = headdim ** -0.5

    def run(self, hq, hk, hv, hdo, ho, hdq=None, hdk=None, hdv=None, causal=False):
        """All h* are pinned host tensors of the full batch; results are written into ho (and hdq/hdk/hdv)."""
        cb, n = self.cb, self.chunks
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_cmp = [torch.cuda.Event() for _ in range(n)]
        ev_out = [torch.cuda.Event() for _ in range(n)]
        for i in range(n):
            sl, bsel = slice(i * cb, (i + 1) * cb), i & 1
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(ev_cmp[i - 2])  # staging buffer bsel is free once chunk i-2 has computed
                self.q[bsel].copy_(hq[sl], non_blocking=True)
                self.k[bsel].copy_(hk[sl], non_blocking=True)
                self.v[bsel].copy_(hv[sl], non_blocking=True)
                if self.bwd:
                    self.do[bsel].copy_(hdo[sl], non_blocking=True)
                ev_in[i].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in[i])
                if i >= 2:
                    self.s_cmp.wait_event(ev_out[i - 2])  # result buffers bsel have been downloaded
                out, lse, _, _ = _op.fwd(self.q[bsel], self.k[bsel], self.v[bsel], self.out[bsel], None, 0.0, self.scale,
                                         causal, -1, -1, 0.0, False, None)
                if self.bwd:
                    _op.bwd(self.do[bsel], self.q[bsel], self.k[bsel], self.v[bsel], out, lse, self.dq[bsel], self.dk[bsel],
                            self.dv[bsel], None, 0.0, self.scale, causal, -1, -1, 0.0, False, None, None)
                ev_cmp[i].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_cmp[i])
                ho[sl].copy_(self.out[bsel], non_blocking=True)
                if self.bwd:
                    hdq[sl].copy_(self.dq[bsel], non_blocking=True)
                    hdk[sl].copy_(self.dk[bsel], non_blocking=True)
                    hdv[sl].copy_(self.dv[bsel], non_blocking=True)
                ev_out[i].record(self.s_out)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_in)

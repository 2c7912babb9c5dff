"""Host-buffer entry point: attention forward (+ backward) on tensors that live in pinned HOST memory.

The C ABI works on device pointers; a caller that owns host buffers (the reference-facing plugin call measured
as `e2e` by bench.py) would otherwise pay H2D -> compute -> D2H back to back.  This helper cuts the problem into
(batch chunk x kv-head group) pieces -- attention is independent per batch element and per kv head, so no result bit
changes -- and runs a 3-stage pipeline on three CUDA streams (H2D of piece i+1 | kernels of piece i | D2H of
piece i-1).  A head group of a `(b, s, h, d)` host tensor is a pitched 2-D region, moved with one
`fa_b200_memcpy2d_async` (cudaMemcpy2DAsync) per tensor into / out of dense device staging buffers.  The step then
costs ~max(H2D, D2H) + one small piece of fill / drain; with 16 pieces at the benchmark shape the PCIe link is busy
in both directions for most of the step.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from . import flash_attn_2_cuda as _op


def bind_to_device_numa_node(device) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, BEFORE it allocates pinned host buffers:
    `cudaHostAlloc` pages are first-touched by the calling thread, so they then land in that node's DRAM and the H2D /
    D2H DMA does not cross the socket interconnect.  Round 1's 8-GPU end-to-end run lost half its per-GPU throughput
    with 8 ranks pinning wherever the kernel put them (GPU0-3 <-> NUMA0, GPU4-7 <-> NUMA1).  Returns what it found;
    a box without the sysfs entries (or a single node) is left untouched."""
    import os

    info = {"numa_node": None, "cpus": None}
    try:
        dev = torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001  (best effort: containers may hide sysfs)
        info["error"] = str(e)[:100]
    return info


class HostAttentionPipeline:
    """Re-usable pipeline (streams, device staging buffers) for one problem shape."""

    def __init__(self, batch, seqlen_q, seqlen_k, nheads, nheads_k, headdim, dtype=torch.bfloat16, device="cuda",
                 chunks: Optional[int] = None, backward=True, head_groups: Optional[int] = None):
        self.dev = torch.device(device)
        self.b, self.bwd = batch, backward
        self.chunks = chunks or batch
        assert batch % self.chunks == 0
        cb = batch // self.chunks
        self.cb = cb
        # kv-head groups per batch chunk: aim at ~16 pieces in total
        if head_groups is None:
            head_groups = 1
            while self.chunks * head_groups < 16 and nheads_k % (head_groups * 2) == 0:
                head_groups *= 2
        assert nheads_k % head_groups == 0
        self.hg = head_groups
        self.hkc = nheads_k // head_groups          # kv heads per piece
        self.hqc = nheads // head_groups            # query heads per piece (whole GQA groups)
        self.h, self.hk, self.d = nheads, nheads_k, headdim
        self.sq, self.sk = seqlen_q, seqlen_k
        self.esz = torch.empty((), dtype=dtype).element_size()
        mk = lambda s, h: [torch.empty(cb, s, h, headdim, dtype=dtype, device=self.dev) for _ in range(2)]
        # double-buffered device staging (piece i+1 is uploaded while piece i computes)
        self.q, self.k, self.v = mk(seqlen_q, self.hqc), mk(seqlen_k, self.hkc), mk(seqlen_k, self.hkc)
        self.do = mk(seqlen_q, self.hqc) if backward else None
        self.out = mk(seqlen_q, self.hqc)
        if backward:
            self.dq, self.dk, self.dv = mk(seqlen_q, self.hqc), mk(seqlen_k, self.hkc), mk(seqlen_k, self.hkc)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.scale = headdim ** -0.5
        self.lib = _lib.load()

    def _copy(self, dev_t, host_t, b0, h0, to_device, stream):
        """dev_t (cb, s, hc, d) dense <-> host_t[b0:b0+cb, :, h0:h0+hc, :] of a dense (b, s, h, d) pinned host tensor."""
        assert host_t.is_contiguous() and dev_t.is_contiguous()
        cb, s, hc, d = dev_t.shape
        h = host_t.shape[2]
        width = hc * d * self.esz
        hpitch = h * d * self.esz
        hptr = host_t.data_ptr() + ((b0 * s) * h + h0) * d * self.esz
        if to_device:
            rc = self.lib.fa_b200_memcpy2d_async(C.c_void_p(dev_t.data_ptr()), width, C.c_void_p(hptr), hpitch, width, cb * s, 1,
                                                 C.c_void_p(stream.cuda_stream))
        else:
            rc = self.lib.fa_b200_memcpy2d_async(C.c_void_p(hptr), hpitch, C.c_void_p(dev_t.data_ptr()), width, width, cb * s, 2,
                                                 C.c_void_p(stream.cuda_stream))
        _lib.check(rc)

    def run(self, hq, hk, hv, hdo, ho, hdq=None, hdk=None, hdv=None, causal=False):
        """All h* are pinned host tensors of the full problem; results are written into ho (and hdq/hdk/hdv)."""
        cb = self.cb
        n = self.chunks * self.hg
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_cmp = [torch.cuda.Event() for _ in range(n)]
        ev_out = [torch.cuda.Event() for _ in range(n)]
        for i in range(n):
            b0, g = (i // self.hg) * cb, i % self.hg
            hq0, hk0 = g * self.hqc, g * self.hkc
            bsel = i & 1
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(ev_cmp[i - 2])  # staging buffer bsel is free once piece i-2 has computed
                self._copy(self.q[bsel], hq, b0, hq0, True, self.s_in)
                self._copy(self.k[bsel], hk, b0, hk0, True, self.s_in)
                self._copy(self.v[bsel], hv, b0, hk0, True, self.s_in)
                if self.bwd:
                    self._copy(self.do[bsel], hdo, b0, hq0, True, self.s_in)
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
                self._copy(self.out[bsel], ho, b0, hq0, False, self.s_out)
                if self.bwd:
                    self._copy(self.dq[bsel], hdq, b0, hq0, False, self.s_out)
                    self._copy(self.dk[bsel], hdk, b0, hk0, False, self.s_out)
                    self._copy(self.dv[bsel], hdv, b0, hk0, False, self.s_out)
                ev_out[i].record(self.s_out)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_in)

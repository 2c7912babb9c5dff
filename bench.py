#!/usr/bin/env python
"""bench.py — headline benchmark of the fused-attention hot path on B200.

Metric (BASELINE.json): attention fwd+bwd TFLOP/s at seqlen=8192 hdim=128, bf16, and the achieved
fraction of the B200 bf16 tensor roofline.  One "step" = one forward + backward pass of
flash_attn_func over one batch of synthetic q/k/v/dO (b=4, h=16, s=8192, d=128, non-causal by default:
BASELINE configs[1]'s shape with the metric's fwd+bwd).  FLOP convention is the reference's
(flash_attn/cute/bench_utils.py:15-48): fwd = 4*b*h*sq*sk*d (/2 causal), bwd = 2.5x, fwd+bwd = 3.5x.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fwd_bwd|fwd] [--causal] [--impl reference]

  value     device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e       the same step through the public host-buffer API (flash_attention_b200.host_pipeline): pinned HOST
            buffers in, H2D of q,k,v,dO and D2H of out,dq,dk,dv inside the timed region (3-stage pipeline over
            the batch: upload | kernels | download)
  roofline  dominant kernel (backward main loop in fwd_bwd mode, forward kernel in fwd mode):
            algorithmic FLOPs / CUDA-event time vs the measured BURST bf16 peak (MEASURED_PEAKS.json; the kernel is
            timed alone); the fraction of the sustained peak is printed beside it
  cpu_baseline / --impl reference: the PyTorch attention_ref restatement (oracle/) on the host cores,
            bounded slice of the same workload (the reference's CPU path; reported, not optimised)

Multi-GPU: the path is per-sequence and does not shard -> N independent replicas (no collective on
the data path); value = N x per-replica work / max-over-ranks time; scaling = weak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORK = dict(b=4, h=16, hk=16, s=8192, d=128)


def flops(b, h, s, d, causal, mode):
    f = 4.0 * b * h * s * s * d / (2 if causal else 1)
    return {"fwd": f, "bwd": 2.5 * f, "fwd_bwd": 3.5 * f}[mode]


def load_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return pk, "measured"
    except Exception:  # noqa: BLE001
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(causal, mode, budget_s=12.0):
    """The reference's CPU path: PyTorch attention_ref restatement (oracle) on the host cores.
    Bounded sample: b=1, h=1, q rows = 1024 (the last rows, bottom-right aligned) x all 8192 keys, d=128."""
    from oracle.attention_oracle import attention_oracle, attention_oracle_bwd

    s, d, rows = WORK["s"], WORK["d"], 1024
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, rows, 1, d, generator=g)
    k = torch.randn(1, s, 1, d, generator=g)
    v = torch.randn(1, s, 1, d, generator=g)
    do = torch.randn(1, rows, 1, d, generator=g)
    pairs = rows * s if not causal else sum(min(s, s - rows + i + 1) for i in range(rows))
    fl = 4.0 * pairs * d * (3.5 if mode == "fwd_bwd" else 1.0)
    fn = (lambda: attention_oracle_bwd(q, k, v, do, causal=causal)) if mode == "fwd_bwd" else (lambda: attention_oracle(q, k, v, causal=causal))
    fn()
    _pick_cpu_threads(fn)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s and n < 50:
        fn()
        n += 1
    dt = (time.perf_counter() - t0) / max(n, 1)
    return {"value": fl / dt / 1e12, "unit": "TFLOP/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"fp32 attention_oracle {mode}, b=1 h=1 last {rows} query rows x {s} keys d={d}, {n} reps, {dt * 1e3:.1f} ms each"}, dt


_CPU_THREADS = None


def _pick_cpu_threads(fn):
    """The oracle is a chain of (1024 x 8192)-sized matmuls: on a 64-core host torch's default (every core) is slower
    than a few threads (measured 0.039 vs 0.156 TFLOP/s).  Calibrate once: time one call at 4, 8, 16, ... threads and
    keep the fastest, so that the CPU arm is the strongest baseline this host gives."""
    global _CPU_THREADS
    if _CPU_THREADS is None:
        ncpu = os.cpu_count() or 1
        cands = sorted({t for t in (4, 8, 16, 32, 64, ncpu) if t <= ncpu} | {min(ncpu, 4)})
        best = (float("inf"), torch.get_num_threads())
        for t in cands:
            torch.set_num_threads(t)
            fn()
            t0 = time.perf_counter()
            fn()
            best = min(best, (time.perf_counter() - t0, t))
        _CPU_THREADS = best[1]
    torch.set_num_threads(_CPU_THREADS)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port), timed per step."""
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1; the CPU arm may use every host core - cpu_baseline() picks the fastest count
    mode = args.mode
    cb, _ = cpu_baseline(args.causal, mode, budget_s=2.0)  # warm
    vals = []
    for _ in range(max(args.warmup, 1)):
        cpu_baseline(args.causal, mode, budget_s=1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c, dt = cpu_baseline(args.causal, mode, budget_s=3.0)
        vals.append(c["value"])
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    v = sum(vals) / len(vals)
    cb["value"] = v
    line = {
        "impl": "reference", "metric": metric_name(mode), "value": v, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(mode):
    return ("attention fwd+bwd TFLOP/s at seqlen=8192 hdim=128; % of B200 bf16 tensor peak" if mode == "fwd_bwd"
            else "attention fwd TFLOP/s at seqlen=8192 hdim=128; % of B200 bf16 tensor peak")


def workload_config(args):
    return {"workload": f"bf16 {'fwd+bwd' if args.mode == 'fwd_bwd' else 'fwd'} {'causal' if args.causal else 'non-causal'} "
                        f"bs={WORK['b']} nheads={WORK['h']} seqlen={WORK['s']} hdim={WORK['d']} per GPU (BASELINE configs[1] shape)",
            "l2": "inputs (q,k,v,dO = 4 x 128 MiB) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"replicas x{args.gpus} (no collective)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", default="fwd_bwd", choices=["fwd_bwd", "fwd"])
    ap.add_argument("--causal", action="store_true")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import flash_attention_b200 as fab
    from flash_attention_b200 import _lib

    lib = _lib.load()  # raises if the CUDA library is missing: no fallback
    b, h, hk, s, d = WORK["b"], WORK["h"], WORK["hk"], WORK["s"], WORK["d"]
    mode, causal = args.mode, args.causal
    torch.manual_seed(rank)
    q = torch.randn(b, s, h, d, device=dev, dtype=torch.bfloat16, requires_grad=(mode == "fwd_bwd"))
    k = torch.randn(b, s, hk, d, device=dev, dtype=torch.bfloat16, requires_grad=(mode == "fwd_bwd"))
    v = torch.randn(b, s, hk, d, device=dev, dtype=torch.bfloat16, requires_grad=(mode == "fwd_bwd"))
    do = torch.randn(b, s, h, d, device=dev, dtype=torch.bfloat16)

    def step():
        out = fab.flash_attn_func(q, k, v, causal=causal)
        if mode == "fwd_bwd":
            out.backward(do)
            q.grad = k.grad = v.grad = None
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("FA_BENCH_NO_SAMPLER"):
        sampler.start()
    launches0 = lib.fa_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = lib.fa_b200_launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    fl = flops(b, h, s, d, causal, mode)
    value = world * fl / (ms_max * 1e-3) / 1e12

    # ---- dominant kernel, timed alone with CUDA events on the launch stream
    from flash_attention_b200 import flash_attn_2_cuda as op
    scale = d ** -0.5
    qd, kd, vd = q.detach(), k.detach(), v.detach()
    out, lse, _, _ = op.fwd(qd, kd, vd, None, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
    if mode == "fwd_bwd":
        dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
        kern = lambda: op.bwd(do, qd, kd, vd, out, lse, dq, dk, dv, None, 0.0, scale, causal, -1, -1, 0.0, False, None, None)
        kern_flops, kern_name = flops(b, h, s, d, causal, "bwd"), "fa_bwd (preprocess + CTA-pair dQ/dK/dV main loop fa_bwd2_kernel + dQ convert)"
    else:
        kern = lambda: op.fwd(qd, kd, vd, out, None, 0.0, scale, causal, -1, -1, 0.0, False, None)
        kern_flops, kern_name = flops(b, h, s, d, causal, "fwd"), "fa_fwd_kernel"
    for _ in range(3):
        kern()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        kern()
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / args.steps
    peaks, peak_src = load_peaks()
    # The dominant kernel is timed ALONE in a sub-second loop, i.e. in the burst regime (SM clock near max): its roofline
    # denominator is the BURST cuBLAS figure (B200_PROFILING.md: "the burst figure for a kernel timed alone, the sustained
    # one for a kernel timed inside a long step"); the fraction against the sustained figure is reported beside it.
    peak = peaks.get("bf16_tflops", 1590.0)
    peak_sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = kern_flops / (kms * 1e-3) / 1e12
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        key = ("bwd" if mode == "fwd_bwd" else "fwd") + ("_causal" if causal else "_noncausal")
        traffic = tr.get(key, {}).get("bytes")  # a number: dram read + write bytes per launch
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "tensor", "kernel": kern_name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": f"bf16_tflops (burst: kernel timed alone), {peak_src}",
                "frac_of_sustained": achieved / peak_sustained, "peak_sustained": peak_sustained,
                "frac_of_nominal_2250": achieved / 2250.0, "ms_per_launch": kms}

    # ---- e2e: host (pinned) buffers -> H2D -> public API -> D2H of out (+ grads)
    from flash_attention_b200.host_pipeline import bind_to_device_numa_node

    numa = bind_to_device_numa_node(dev)   # pinned buffers below are first-touched on the GPU's own NUMA node
    hq, hk_, hv, hdo = (torch.empty(t_.shape, dtype=t_.dtype).pin_memory() for t_ in (qd, kd, vd, do))
    for dst, src in ((hq, qd), (hk_, kd), (hv, vd), (hdo, do)):
        dst.copy_(src)
    ho = torch.empty_like(hq).pin_memory()  # results land in pinned host memory too (pageable D2H is synchronous)
    hdq, hdk, hdv = torch.empty_like(hq).pin_memory(), torch.empty_like(hk_).pin_memory(), torch.empty_like(hv).pin_memory()

    from flash_attention_b200.host_pipeline import HostAttentionPipeline

    # public host-buffer entry point: 3-stage pipeline over the batch (H2D | kernels | D2H on separate streams)
    pipe = HostAttentionPipeline(b, s, s, h, hk, d, torch.bfloat16, dev, chunks=b, backward=(mode == "fwd_bwd"))

    def e2e_step():
        pipe.run(hq, hk_, hv, hdo, ho, hdq, hdk, hdv, causal=causal)

    nbytes = lambda t_: t_.numel() * t_.element_size()
    h2d = nbytes(hq) + nbytes(hk_) + nbytes(hv) + (nbytes(hdo) if mode == "fwd_bwd" else 0)
    d2h = nbytes(ho) + ((nbytes(hdq) + nbytes(hdk) + nbytes(hdv)) if mode == "fwd_bwd" else 0)
    for _ in range(2):
        e2e_step()
    barrier()
    esteps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    for _ in range(esteps):
        e2e_step()
    x1.record()
    barrier()
    ems = torch.tensor([x0.elapsed_time(x1) / esteps], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e = {"value": world * fl / (ems.item() * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ems.item(), "numa": numa}

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:   # reported at N = 1 only
            cb, _ = cpu_baseline(causal, mode)
        line = {
            "metric": metric_name(mode), "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "pct_of_bf16_peak": {"sustained_measured": 100 * value / world / peaks.get("bf16_tflops_sustained", 1400.0),
                                 "burst_measured": 100 * value / world / peaks.get("bf16_tflops", 1590.0),
                                 "nominal_2250": 100 * value / world / 2250.0},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep.py ours > gpurun_out/sweep_ours_v7.jsonl 2>/dev/null
timeout 900 python tools/sweep.py fa2 > gpurun_out/sweep_fa2_v7.jsonl 2>/dev/null
wc -l gpurun_out/sweep_ours_v7.jsonl gpurun_out/sweep_fa2_v7.jsonl

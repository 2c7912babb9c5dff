#!/bin/bash
mkdir -p gpurun_out
timeout 60 experiments/mma_rate_probe.bin 2>&1 | tee gpurun_out/mma_rate2.txt
H=$PWD/flash_attention_b200
export FA_B200_BWD_PAIR=1
export FA_B200_LIB=$H/libfa_b200_ts.so
timeout 120 python tools/probe_bwd_ts.py nc brief 2>&1 | tail -24 | tee gpurun_out/y_ts.txt

#!/bin/bash
mkdir -p gpurun_out
H=$PWD/flash_attention_b200
export FA_B200_BWD_PAIR=1
export FA_B200_LIB=$H/libfa_b200_ts.so
timeout 120 python tools/probe_bwd_ts.py nc 2>&1 | tail -24 | tee gpurun_out/y_ts.txt
export FA_B200_LIB=$H/libfa_b200.so
timeout 100 python tools/probe_bwd.py small 2>&1 | grep -E "OK|FAIL|timeout|rror" | head -4 | cut -c1-200
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"

#!/bin/bash
mkdir -p gpurun_out
H=$PWD/flash_attention_b200
export FA_B200_BWD_PAIR=1
export FA_B200_LIB=$H/libfa_b200_guard.so
timeout 150 python tools/probe_bwd.py small 2>&1 | grep -E "OK|FAIL|timeout|rror" | cut -c1-200
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "small stage failed/hung rc=$rc; stop"; exit 0; fi
timeout 150 python tools/probe_bwd.py masks 2>&1 | grep -E "FAIL|timeout|rror|done" | cut -c1-200
timeout 150 python tools/probe_bwd.py gqa 2>&1 | grep -E "FAIL|timeout|rror|done" | cut -c1-200
export FA_B200_LIB=$H/libfa_b200.so
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
export FA_B200_LIB=$H/libfa_b200_ts.so
timeout 120 python tools/probe_bwd_ts.py nc 2>&1 | tail -24 > gpurun_out/y_ts.txt; sed -n 3,6p gpurun_out/y_ts.txt | cut -c1-700
export FA_B200_LIB=$H/libfa_b200.so
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py -q -m gpu -x 2>&1 | tail -4

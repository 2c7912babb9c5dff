#!/bin/bash
H=$PWD/flash_attention_b200
export FA_B200_LIB=$H/libfa_b200_guard.so
timeout 150 python tools/probe_bwd.py small 2>&1 | grep -E "OK|FAIL|timeout|rror" | cut -c1-120
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "small stage failed/hung rc=$rc; stop"; exit 0; fi
export FA_B200_LIB=$H/libfa_b200_ts.so
timeout 120 python tools/probe_bwd_ts.py nc brief 2>&1 | grep -E "once" | cut -c1-400
unset FA_B200_LIB
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py tests/test_dropin_gpu.py tests/test_bwd_variants_gpu.py -q -m gpu 2>&1 | tail -3
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"

#!/bin/bash
mkdir -p gpurun_out
H=$PWD/flash_attention_b200
export FA_B200_BWD_PAIR=1
export FA_B200_LIB=$H/libfa_b200_guard.so
echo "== pair kernel (guard build): small + masks"
timeout 150 python tools/probe_bwd.py small 2>&1 | grep -E "OK|FAIL|timeout|rror" | cut -c1-200
rc=${PIPESTATUS[0]}
echo "rc=$rc"
if [ "$rc" != "0" ]; then echo "small stage failed/hung; stop"; exit 0; fi
timeout 150 python tools/probe_bwd.py masks 2>&1 | grep -E "OK|FAIL|timeout|rror" | cut -c1-200
timeout 150 python tools/probe_bwd.py gqa 2>&1 | grep -E "OK|FAIL|timeout|rror" | cut -c1-200
for v in ""; do
  export FA_B200_LIB=$H/libfa_b200$v.so
  echo "== time pair variant '$v'"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done
export FA_B200_LIB=$H/libfa_b200.so
echo "== time single"
FA_B200_BWD_PAIR=0 timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
echo "== pytest bwd (pair)"
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py -q -m gpu -x 2>&1 | tail -4

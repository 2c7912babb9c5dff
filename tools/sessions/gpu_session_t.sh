#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
for v in "" _dq16; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 100 python tools/probe_bwd.py small 2>&1 | grep -E "multi|ragged|fp16|FAIL" | cut -c1-150
  timeout 100 python tools/probe_bwd.py masks 2>&1 | grep -E "causal\]|FAIL" | cut -c1-150
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done
export FA_B200_LIB=$PWD/$H/libfa_b200_dq16.so
timeout 300 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py -q -m gpu 2>&1 | tail -5

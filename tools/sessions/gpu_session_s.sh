#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 300 python -m pytest tests/test_bwd_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -4
unset FA_B200_LIB
for v in "" _v3; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done

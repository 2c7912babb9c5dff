#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/d_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -25 gpurun_out/d_pytest_guard.log
timeout 120 python tools/probe_fwd.py fp8 > gpurun_out/d_fp8.log 2>&1; echo "fp8 exit $?"; grep -E "fp8|time|rror" gpurun_out/d_fp8.log | head -20
unset FA_B200_LIB
for v in "" _poly0 _poly1 _poly3 _poly4; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]" | head -2
done

#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 300 python -m pytest tests/test_bwd_gpu.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/g_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -12 gpurun_out/g_pytest_guard.log
timeout 100 python tools/probe_bwd.py masks > gpurun_out/g_bwd_masks.log 2>&1; grep -E "OK|FAIL|rror|timeout" gpurun_out/g_bwd_masks.log | head -12
unset FA_B200_LIB
for v in "" _bpoly0 _bpoly2; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done

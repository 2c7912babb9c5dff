#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 200 python -m pytest tests/test_bwd_gpu.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/i_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -6 gpurun_out/i_pytest_guard.log
timeout 100 python tools/probe_bwd.py masks > gpurun_out/i_bwd_masks.log 2>&1; grep -E "OK|FAIL|rror|timeout" gpurun_out/i_bwd_masks.log | cut -c1-60 | head -12
timeout 100 python tools/probe_bwd.py gqa 2>&1 | grep -E "OK|FAIL|rror|timeout" | cut -c1-60
unset FA_B200_LIB
for v in "" _cw2 _bpoly0; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done

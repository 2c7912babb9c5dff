#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 60 python tools/probe_fwd.py s_tile 2>&1 | grep -E "S tile|O \(|LSE|rror|timeout" 
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/k_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -8 gpurun_out/k_pytest_guard.log
unset FA_B200_LIB
for v in "" _v2 _poly0 _poly2; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 100 python tools/probe_fwd.py time 2>&1 | grep "time\]"
done
unset FA_B200_LIB
timeout 100 python tools/probe_fwd.py fp8 2>&1 | grep "C4"

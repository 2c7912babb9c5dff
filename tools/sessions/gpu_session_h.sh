#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
timeout 120 python tools/probe_bwd_scale.py 2>&1 | grep bwd-scale
for v in "" _red32 _red64 _red128; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 100 python tools/probe_bwd.py small 2>&1 | grep -E "1blk|multi|FAIL" | head -3
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]" | head -1
done
unset FA_B200_LIB
timeout 120 python tools/probe_fwd.py fp8 2>&1 | grep -E "fp8|C4"

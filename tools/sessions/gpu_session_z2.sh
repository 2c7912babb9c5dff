#!/bin/bash
mkdir -p gpurun_out
H=$PWD/flash_attention_b200
export FA_B200_BWD_PAIR=1
for v in "" _localds; do
  export FA_B200_LIB=$H/libfa_b200$v.so
  echo "== time pair variant '$v'"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]" | head -1
done

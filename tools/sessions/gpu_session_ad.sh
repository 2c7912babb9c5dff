#!/bin/bash
H=$PWD/flash_attention_b200
for v in "" _rot1 _rot5; do
  export FA_B200_LIB=$H/libfa_b200$v.so
  echo "== variant '$v'"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done
export FA_B200_LIB=$H/libfa_b200_rot5.so
timeout 300 python -m pytest tests/test_bwd_gpu.py -q -m gpu 2>&1 | tail -2
unset FA_B200_LIB
timeout 600 python -m pytest tests/test_bwd_variants_gpu.py -q -m gpu 2>&1 | tail -2

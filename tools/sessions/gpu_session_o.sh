#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
timeout 300 python -m pytest tests/test_fa3_surface_gpu.py tests/test_kvcache_gpu.py tests/test_fwd_gpu.py -x -q -m gpu 2>&1 | tail -5
for v in "" _poly2 _poly3; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 100 python tools/probe_d64.py 2>&1 | grep "d64 probe"
done

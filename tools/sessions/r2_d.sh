#!/bin/bash
# round 2, GPU call D: persistent forward -- first under the spin guard (a starved barrier traps instead of hanging)
mkdir -p gpurun_out
echo "== guard build: forward tests"
FA_B200_LIB=flash_attention_b200/variants/lib_guard.so timeout 600 python -m pytest tests/test_fwd_gpu.py tests/test_kvcache_gpu.py tests/test_fp8_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2d_guard.log
if grep -q "failed\|error\|Timeout\|timeout" gpurun_out/r2d_guard.log; then echo "GUARD RUN FAILED - stopping"; exit 0; fi
echo "== default build: all forward-side + drop-in + custom-op tests"
timeout 1500 python -m pytest tests -q -m gpu -x -k "not bwd" 2>&1 | tail -15 | tee gpurun_out/r2d_pytest.log
echo "== timing"
for cfg in "default:" "causal1tile:FA_B200_FWD_TILE_LIMIT_CAUSAL=64" "group4:FA_B200_SCHED_GROUP=4" "group64:FA_B200_SCHED_GROUP=64"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs FA_TAG=$tag timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -8
done | tee gpurun_out/r2d_variants.log
FA_TAG=p3 FA_B200_LIB=flash_attention_b200/variants/lib_p3.so timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -8 | tee -a gpurun_out/r2d_variants.log
timeout 300 python tools/bench_configs.py 2>&1 | tail -9 | tee gpurun_out/r2d_configs.txt | cut -c1-200

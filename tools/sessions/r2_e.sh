#!/bin/bash
# round 2, GPU call E: persistent CTA-pair backward -- under the spin guard first, then the full suite and A/B timing
mkdir -p gpurun_out
echo "== guard build: backward tests"
FA_B200_LIB=flash_attention_b200/variants/lib_guard.so timeout 900 python -m pytest tests/test_bwd_gpu.py tests/test_bwd_stress_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2e_guard.log
if grep -q "failed\|error\|Timeout\|timeout" gpurun_out/r2e_guard.log; then echo "GUARD RUN FAILED - stopping"; exit 0; fi
echo "== default build: full suite"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2e_pytest.log
echo "== A/B timing, same lease"
for v in old default p3 r216; do
  if [ $v = default ]; then unset FA_B200_LIB; else export FA_B200_LIB=flash_attention_b200/variants/lib_$v.so; fi
  FA_TAG=$v timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -8
  echo "-- fwd+bwd ($v)"; timeout 300 python tools/baseline_probe.py ours 2>&1 | grep "fwd+bwd\|bwd(diff)" | head -10
done | tee gpurun_out/r2e_ab.log
unset FA_B200_LIB
for cfg in "causal1tile:FA_B200_FWD_TILE_LIMIT_CAUSAL=64" "group2:FA_B200_SCHED_GROUP=2" "group32:FA_B200_SCHED_GROUP=32" "bgroup1:FA_B200_SCHED_GROUP_BWD=1" "bgroup16:FA_B200_SCHED_GROUP_BWD=16"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs FA_TAG=$tag timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -1
  env $envs timeout 300 python tools/baseline_probe.py ours 2>&1 | grep "bwd(diff)" | head -3
done | tee gpurun_out/r2e_knobs.log
timeout 300 python bench.py 2>gpurun_out/r2e_bench.err | tail -1 | tee gpurun_out/r2e_bench.json | cut -c1-1200

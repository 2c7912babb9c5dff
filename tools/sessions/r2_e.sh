#!/bin/bash
# round 2, GPU call E: persistent CTA-pair backward -- under the spin guard first, then the full suite and timing
mkdir -p gpurun_out
echo "== guard build: backward tests"
FA_B200_LIB=flash_attention_b200/variants/lib_guard.so timeout 900 python -m pytest tests/test_bwd_gpu.py tests/test_bwd_stress_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2e_guard.log
if grep -q "failed\|error\|Timeout\|timeout" gpurun_out/r2e_guard.log; then echo "GUARD RUN FAILED - stopping"; exit 0; fi
echo "== default build: full suite"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2e_pytest.log
echo "== timing"
timeout 300 python tools/baseline_probe.py ours 2>&1 | grep -v Warning | tail -20 | tee gpurun_out/r2e_probe_ours.log
for g in 1 4 64; do
  echo "-- FA_B200_SCHED_GROUP_BWD=$g"; FA_B200_SCHED_GROUP_BWD=$g timeout 300 python tools/baseline_probe.py ours 2>&1 | grep "fwd+bwd\|bwd(diff)" | head -6
done | tee gpurun_out/r2e_groups.log
timeout 300 python bench.py 2>gpurun_out/r2e_bench.err | tail -1 | tee gpurun_out/r2e_bench.json | cut -c1-1200

#!/bin/bash
# round 2, GPU call H: one-pass softmax blocks (stale reference + row-sum guard): guard build first, suite, ABAB vs the two-pass build
mkdir -p gpurun_out
echo "== guard build: forward tests"
FA_B200_LIB=flash_attention_b200/variants/lib_guard.so timeout 1500 python -m pytest tests/test_fwd_onepass_gpu.py tests/test_fwd_gpu.py tests/test_sched_dynamic_gpu.py tests/test_fwd_stress_gpu.py tests/test_reference_matrix_gpu.py tests/test_varlen_matrix_gpu.py tests/test_kvcache_gpu.py tests/test_fa2_fullsize_gpu.py tests/test_baseline_configs_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2h_guard.log
if grep -q " failed\|error\|Timeout\|timeout" gpurun_out/r2h_guard.log; then echo "GUARD RUN FAILED"; FAILED=1; fi
if [ -z "$FAILED" ]; then
echo "== default build: full suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2h_pytest.log
fi
echo "== ABAB timing one-pass vs two-pass vs r200"
for rep in 1 2; do for v in default nopass r200; do
  if [ $v = default ]; then unset FA_B200_LIB; else export FA_B200_LIB=flash_attention_b200/variants/lib_$v.so; fi
  FA_TAG=$v timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -1
done; done | tee gpurun_out/r2h_abab.log
unset FA_B200_LIB
echo "== timeline"
(FA_B200_LIB=flash_attention_b200/variants/lib_ts.so timeout 300 python tools/probe_fwd_ts.py; FA_B200_LIB=flash_attention_b200/variants/lib_ts.so timeout 300 python tools/probe_fwd_items.py causal) 2>&1 | grep -v Warn | tee gpurun_out/r2h_timeline.log | cut -c1-300
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warn | tee gpurun_out/r2h_configs.txt | cut -c1-200
timeout 600 python tools/baseline_probe.py ours 2>&1 | grep -v Warn | tee gpurun_out/r2h_ours.log | cut -c1-200

#!/bin/bash
# round 2, GPU call G: dynamic work-item scheduler (forward + pair backward): guard build first, then suite, ABAB timing
mkdir -p gpurun_out
echo "== guard build: forward / backward / varlen / kvcache tests"
FA_B200_LIB=flash_attention_b200/variants/lib_guard.so timeout 1500 python -m pytest tests/test_fwd_gpu.py tests/test_bwd_gpu.py tests/test_bwd_stress_gpu.py tests/test_varlen_matrix_gpu.py tests/test_sched_dynamic_gpu.py tests/test_kvcache_gpu.py tests/test_fa2_fullsize_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2g_guard.log
if grep -q " failed\|error\|Timeout\|timeout" gpurun_out/r2g_guard.log; then echo "GUARD RUN FAILED - stopping"; exit 0; fi
echo "== default build: full suite"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2g_pytest.log
echo "== ABAB timing dynamic vs static"
for rep in 1 2; do for dynv in 1 0; do
  FA_B200_SCHED_DYNAMIC=$dynv FA_TAG=dyn$dynv timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -1
  FA_B200_SCHED_DYNAMIC=$dynv timeout 300 python tools/baseline_probe.py ours 2>&1 | grep "bwd(diff)" | head -5
done; done | tee gpurun_out/r2g_abab.log
echo "== item timeline (dynamic)"
(FA_B200_LIB=flash_attention_b200/variants/lib_ts.so timeout 300 python tools/probe_fwd_items.py; FA_B200_LIB=flash_attention_b200/variants/lib_ts.so timeout 300 python tools/probe_fwd_items.py causal) 2>&1 | grep -v Warn | tee gpurun_out/r2g_items.log
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warn | tee gpurun_out/r2g_configs.txt | cut -c1-200

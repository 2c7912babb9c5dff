#!/bin/bash
# round 2, GPU call F: cleaner-race fix (kvcache matrix x3), dropout stream vs FA2, drop-in/custom-op tests, per-item timeline
mkdir -p gpurun_out
echo "== kvcache matrix x3 + split tests"
for i in 1 2 3; do timeout 600 python -m pytest tests/test_kvcache_matrix_gpu.py tests/test_kvcache_gpu.py -x -q 2>&1 | tail -3; done | tee gpurun_out/r2f_kvcache.log
echo "== dropout / drop-in / custom ops / fa2 fullsize"
timeout 1200 python -m pytest tests/test_dropout_gpu.py tests/test_dropin_versions_gpu.py tests/test_custom_ops_gpu.py tests/test_fa2_fullsize_gpu.py tests/test_dropin_gpu.py -q 2>&1 | tail -25 | tee gpurun_out/r2f_new.log
echo "== item timeline"
(FA_B200_LIB=flash_attention_b200/variants/lib_ts.so timeout 300 python tools/probe_fwd_items.py; FA_B200_LIB=flash_attention_b200/variants/lib_ts.so timeout 300 python tools/probe_fwd_items.py causal) 2>&1 | grep -v Warn | tee gpurun_out/r2f_items.log
echo "== full suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2f_pytest.log
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warn | tee gpurun_out/r2f_configs.txt | cut -c1-200

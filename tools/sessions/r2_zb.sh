#!/bin/bash
# round 2, GPU call ZB: forward-facing test files after the kernel-choice change (2-tile from 5 blocks per item on)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fwd_gpu.py tests/test_fwd_stress_gpu.py tests/test_reference_matrix_gpu.py tests/test_varlen_matrix_gpu.py tests/test_kvcache_gpu.py tests/test_kvcache_matrix_gpu.py tests/test_sched_dynamic_gpu.py tests/test_fp8_gpu.py tests/test_fa2_fullsize_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py tests/test_custom_ops_gpu.py tests/test_fa3_surface_gpu.py -q 2>&1 | tail -4 | tee gpurun_out/r2zb_pytest.log
timeout 300 python tools/stress_fwd.py 31 40 2>&1 | tail -1 | tee -a gpurun_out/r2zb_pytest.log

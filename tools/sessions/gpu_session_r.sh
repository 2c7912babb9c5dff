#!/bin/bash
mkdir -p gpurun_out
export FA_B200_LIB=$PWD/flash_attention_b200/libfa_b200_guard.so
timeout 600 python -m pytest tests/test_fwd_gpu.py tests/test_kvcache_gpu.py -x -q -m gpu > gpurun_out/r_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -12 gpurun_out/r_pytest_guard.log | cut -c1-200

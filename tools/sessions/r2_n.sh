#!/bin/bash
# round 2, GPU call N: first run of the wide-S forward kernel (N = 256): tests (spin-guard build first), then timing against the two-tile kernel
mkdir -p gpurun_out
FA_B200_LIB=flash_attention_b200/variants/lib_wguard.so timeout 600 python -m pytest tests/test_fwd_wide_gpu.py -x -q 2>&1 | tail -25 | tee gpurun_out/r2n_wide_tests.log
if grep -q " failed\|error\|Timeout\|timeout\|Error" gpurun_out/r2n_wide_tests.log; then echo "GUARD RUN FAILED"; exit 0; fi
for w in 0 1; do FA_B200_FWD_WIDE=$w FA_TAG=wide$w timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warn | tail -8; done | tee gpurun_out/r2n_quick.log

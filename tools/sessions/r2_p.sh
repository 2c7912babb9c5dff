#!/bin/bash
# round 2, GPU call P: wide-S kernel with partner-warp exchange + streaming S load; polynomial share 3/8
mkdir -p gpurun_out
V=flash_attention_b200/variants
FA_B200_LIB=$V/lib_wguard.so timeout 600 python -m pytest tests/test_fwd_wide_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2p_wide_tests.log
if grep -q " failed\|error\|Timeout\|timeout\|Error" gpurun_out/r2p_wide_tests.log; then echo "GUARD RUN FAILED"; exit 0; fi
(FA_B200_FWD_WIDE=0 FA_TAG=two-tile timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide timeout 200 python tools/probe_fwd_wide.py
 FA_B200_LIB=$V/lib_wp3.so FA_B200_FWD_WIDE=1 FA_TAG=wide-poly3 timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide timeout 200 python tools/probe_fwd_wide.py) 2>&1 | grep -v Warn | tee gpurun_out/r2p_timing.log

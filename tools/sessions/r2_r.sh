#!/bin/bash
# round 2, GPU call R: wide-S kernel with four softmax groups (16 warps): tests, timing vs two groups vs the two-tile kernel, timeline
mkdir -p gpurun_out
V=flash_attention_b200/variants
FA_B200_LIB=$V/lib_wguard.so timeout 600 python -m pytest tests/test_fwd_wide_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2r_wide_tests.log
if grep -q " failed\|error\|Timeout\|timeout\|Error" gpurun_out/r2r_wide_tests.log; then echo "GUARD RUN FAILED"; exit 0; fi
(FA_B200_FWD_WIDE=0 FA_TAG=two-tile timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide4 timeout 200 python tools/probe_fwd_wide.py
 FA_B200_LIB=$V/lib_wg2.so FA_B200_FWD_WIDE=1 FA_TAG=wide2 timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide4 timeout 200 python tools/probe_fwd_wide.py) 2>&1 | grep -v Warn | tee gpurun_out/r2r_timing.log
FA_B200_LIB=$V/lib_wts.so timeout 200 python tools/probe_fwd_wide_ts.py 2>&1 | grep -v Warn | tee gpurun_out/r2r_wide_timeline.log | cut -c1-400

#!/bin/bash
# round 2, GPU call U: wide-S kernel, one-pass block pairs (lagging reference, exchange after the hand-over): tests (guard build twice +
# default build), timing, timeline
mkdir -p gpurun_out
V=flash_attention_b200/variants
for rep in 1 2; do FA_B200_LIB=$V/lib_wguard.so timeout 900 python -m pytest tests/test_fwd_wide_gpu.py -x -q 2>&1 | tail -6; done | tee gpurun_out/r2u_wide_tests.log
if grep -q " failed\|error\|Timeout\|timeout\|Error" gpurun_out/r2u_wide_tests.log; then echo "GUARD RUN FAILED"; exit 0; fi
timeout 900 python -m pytest tests/test_fwd_wide_gpu.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2u_wide_tests.log
(FA_B200_FWD_WIDE=0 FA_TAG=two-tile timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide-onepass timeout 200 python tools/probe_fwd_wide.py
 FA_B200_LIB=$V/lib_wtp.so FA_B200_FWD_WIDE=1 FA_TAG=wide-twopass timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide-onepass timeout 200 python tools/probe_fwd_wide.py) 2>&1 | grep -v Warn | tee gpurun_out/r2u_timing.log
FA_B200_LIB=$V/lib_wts.so timeout 200 python tools/probe_fwd_wide_ts.py 2>&1 | grep -v Warn | tee gpurun_out/r2u_wide_timeline.log | cut -c1-400

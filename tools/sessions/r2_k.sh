#!/bin/bash
# round 2, GPU call K: OR-guard one-pass (late row sum) vs two-pass late sum vs HEAD; tests on the OR-guard build; timelines
mkdir -p gpurun_out
V=flash_attention_b200/variants
FA_B200_LIB=$V/lib_og.so timeout 900 python -m pytest tests/test_fwd_onepass_gpu.py tests/test_fwd_gpu.py tests/test_fa2_fullsize_gpu.py tests/test_fwd_stress_gpu.py tests/test_sched_dynamic_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2k_og_tests.log
for rep in 1 2; do for v in default sl og; do
  if [ $v = default ]; then unset FA_B200_LIB; else export FA_B200_LIB=$V/lib_$v.so; fi
  FA_TAG=$v timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warn | tail -1
done; done | tee gpurun_out/r2k_abab.log
for v in slts ogts; do echo "== timeline $v"; FA_B200_LIB=$V/lib_$v.so timeout 300 python tools/probe_fwd_ts.py 2>&1 | grep -v Warn; done | tee gpurun_out/r2k_timeline.log | cut -c1-600

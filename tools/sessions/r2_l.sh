#!/bin/bash
# round 2, GPU call L: OR-guard one-pass v2 (no held-back stores) vs two-pass late sum; never-firing-guard build as a timing bound
mkdir -p gpurun_out
V=flash_attention_b200/variants
FA_B200_LIB=$V/lib_og.so timeout 900 python -m pytest tests/test_fwd_onepass_gpu.py tests/test_fwd_gpu.py tests/test_fa2_fullsize_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2l_og_tests.log
for rep in 1 2; do for v in sl og ognf; do
  export FA_B200_LIB=$V/lib_$v.so
  FA_TAG=$v timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warn | tail -1
done; done | tee gpurun_out/r2l_abab.log

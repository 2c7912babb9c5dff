#!/bin/bash
# round 2, GPU call J: forward softmax-chain variants (late row sum, 3/4 + 1/4 P hand-over, one-pass on/off), ABAB + timelines
mkdir -p gpurun_out
V=flash_attention_b200/variants
for rep in 1 2; do for v in default nopass sl sl96 np96; do
  if [ $v = default ]; then unset FA_B200_LIB; else export FA_B200_LIB=$V/lib_$v.so; fi
  FA_TAG=$v timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warn | tail -1
done; done | tee gpurun_out/r2j_abab.log
for v in ts sl96ts; do echo "== timeline $v"; FA_B200_LIB=$V/lib_$v.so timeout 300 python tools/probe_fwd_ts.py 2>&1 | grep -v Warn; done | tee gpurun_out/r2j_timeline.log | cut -c1-400
FA_B200_LIB=$V/lib_sl96.so timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_fp8_gpu.py tests/test_fa2_fullsize_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2j_sl96_tests.log

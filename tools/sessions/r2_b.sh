#!/bin/bash
# round 2, GPU call B: forward exp-phase variants A/B + timelines + ex77 + ncu exports (small files only)
mkdir -p gpurun_out
for v in r1 p2 p3 p4; do
  FA_TAG=$v FA_B200_LIB=flash_attention_b200/variants/lib_$v.so timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warning | tail -9
done | tee gpurun_out/r2b_variants.log
for v in tsr1 ts; do
  echo "== timeline $v"; FA_B200_LIB=flash_attention_b200/variants/lib_$v.so timeout 200 python tools/probe_fwd_ts.py 2>&1 | tail -8
done | tee gpurun_out/r2b_timeline.log
timeout 600 python -m pytest tests/test_fwd_gpu.py tests/test_fp8_gpu.py tests/test_kvcache_gpu.py tests/test_fa2_fullsize_gpu.py -q -x 2>&1 | tail -4 | tee gpurun_out/r2b_pytest.log
if [ -x baseline/_ref/ex77/77_blackwell_fmha_fp16 ]; then
  for m in no causal; do
    timeout 120 baseline/_ref/ex77/77_blackwell_fmha_fp16 --b=4 --h=16 --q=8192 --k=8192 --d=128 --mask=$m --iterations=10 2>&1 | tail -4
    timeout 120 baseline/_ref/ex77/77_blackwell_fmha_bwd_fp16 --b=4 --h=16 --q=8192 --k=8192 --d=128 --mask=$m --iterations=10 2>&1 | tail -4
  done | tee gpurun_out/r2b_ex77.log
fi
for spec in fp8_c4:4:3:1:r2b_fp8c4 bf16_c4:4:3:1:r2b_bf16c4 decode_c5:4:6:2:r2b_c5 bwd_hd256:3:8:4:r2b_hd256 hd64:3:8:4:r2b_hd64 headline:3:8:4:r2b_headline headline_causal:3:8:4:r2b_causal; do
  IFS=: read c reps skip cnt tag <<< "$spec"
  bash tools/ncu_export.sh $c $reps $skip $cnt $tag
done
ls -la gpurun_out | head -40

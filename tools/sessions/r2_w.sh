#!/bin/bash
# round 2, GPU call W: polynomial share of the two-tile forward re-tuned with the late row sum (1 / 2 / 3 of 8 pairs), ABAB
mkdir -p gpurun_out
V=flash_attention_b200/variants
for rep in 1 2; do for v in default p3 p1; do
  if [ $v = default ]; then unset FA_B200_LIB; else export FA_B200_LIB=$V/lib_$v.so; fi
  FA_TAG=$v timeout 300 python tools/probe_fwd_quick.py 2>&1 | grep -v Warn | tail -1
done; done | tee gpurun_out/r2w_poly_abab.log

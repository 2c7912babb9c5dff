#!/bin/bash
# round 2, GPU call O: where the wide-S kernel's time goes (timing-only variants: half the exponentials, no max exchange, L2-resident K/V)
mkdir -p gpurun_out
V=flash_attention_b200/variants
(FA_B200_FWD_WIDE=0 FA_TAG=two-tile timeout 200 python tools/probe_fwd_wide.py
 FA_B200_FWD_WIDE=1 FA_TAG=wide timeout 200 python tools/probe_fwd_wide.py
 for v in wd1 wd2 wd3; do FA_B200_LIB=$V/lib_$v.so FA_B200_FWD_WIDE=1 FA_TAG=$v timeout 200 python tools/probe_fwd_wide.py; done) 2>&1 | grep -v Warn | tee gpurun_out/r2o_diag.log

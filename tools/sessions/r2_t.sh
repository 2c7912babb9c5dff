#!/bin/bash
# round 2, GPU call T: ncu evidence of the final round-2 build: launch list of the bench command, full-set captures of the headline
# forward (two-tile, late row sum) + CTA-pair backward, and of the wide-S forward kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2t_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r2t_bench_under_ncu.log | cut -c1-200
# skip the 2 warm-up iterations' 8 launches (fwd, preprocess, bwd2, convert per iteration), capture one iteration
bash tools/ncu_export.sh headline 3 8 4 r2t_headline
FA_B200_FWD_WIDE=1 bash tools/ncu_export.sh headline_fwd 3 2 1 r2t_wide_fwd
bash tools/ncu_export.sh headline_causal 3 8 4 r2t_causal
ls -la gpurun_out | grep r2t

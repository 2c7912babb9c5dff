#!/bin/bash
mkdir -p gpurun_out
export FA_B200_BWD_PAIR=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_bwd2_kernel" -s 3 -c 1 -o gpurun_out/prof_r01_pair -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/w_ncu_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/w_ncu_full.log | cut -c1-300

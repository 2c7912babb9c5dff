#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_(fwd|bwd2)_kernel" -s 6 -c 2 -o gpurun_out/prof_r01_v7_causal -f python bench.py --causal --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/az_ncu_causal.log 2>&1; echo "ncu causal exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_(fwd|bwd2)_kernel" -s 6 -c 2 -o gpurun_out/prof_r01_v7 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/az_ncu.log 2>&1; echo "ncu exit $?"

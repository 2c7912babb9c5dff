#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -8 gpurun_out/e_pytest_guard.log
timeout 120 python tools/probe_fwd.py fp8 > gpurun_out/e_fp8.log 2>&1; echo "fp8 exit $?"; grep -E "fp8|time|rror" gpurun_out/e_fp8.log | head -20
unset FA_B200_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_(fwd|bwd)_kernel" -s 6 -c 2 -o gpurun_out/prof_r01_v2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/e_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; cat gpurun_out/e_bench.json | cut -c1-400

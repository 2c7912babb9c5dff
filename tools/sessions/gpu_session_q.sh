#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/q_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -5 gpurun_out/q_pytest_guard.log
unset FA_B200_LIB
timeout 400 python tools/bench_configs.py ours > gpurun_out/q_configs.jsonl 2>/dev/null; cat gpurun_out/q_configs.jsonl | cut -c1-170
timeout 600 python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; cut -c1-200 gpurun_out/q_bench.json
timeout 600 python bench.py --causal --no-cpu-baseline > gpurun_out/q_bench_causal.json 2>> gpurun_out/q_bench.err; cut -c1-200 gpurun_out/q_bench_causal.json

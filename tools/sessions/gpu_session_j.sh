#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/j_pytest_guard.log 2>&1; echo "pytest(guard) exit $?"; tail -8 gpurun_out/j_pytest_guard.log
unset FA_B200_LIB
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; cat gpurun_out/j_bench.json | cut -c1-300
timeout 600 python bench.py --causal --no-cpu-baseline > gpurun_out/j_bench_causal.json 2>> gpurun_out/j_bench.err; cat gpurun_out/j_bench_causal.json | cut -c1-200

#!/bin/bash
# round-end validation of the committed default build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py 2>gpurun_out/final_bench.err | tail -1 | tee gpurun_out/final_bench.json | cut -c1-1200
timeout 300 python bench.py --causal --no-cpu-baseline 2>>gpurun_out/final_bench.err | tail -1 | tee gpurun_out/final_bench_causal.json | cut -c1-300
timeout 300 python bench.py --mode fwd --no-cpu-baseline 2>>gpurun_out/final_bench.err | tail -1 | tee gpurun_out/final_bench_fwd.json | cut -c1-300
timeout 300 python tools/bench_configs.py 2>&1 | tail -8 | tee gpurun_out/final_configs.txt | cut -c1-200

#!/bin/bash
# round 2, GPU call I (re-entry): state of HEAD -- full GPU suite, smoke, bench lines, ours vs FA4(shim) on one lease
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/r2i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/r2i_smoke.log
timeout 600 python bench.py 2>&1 | grep -v Warn | tail -2 | tee gpurun_out/r2i_bench.json
timeout 300 python bench.py --mode fwd 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/r2i_bench_fwd.json
timeout 300 python bench.py --causal 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/r2i_bench_causal.json
timeout 600 python tools/baseline_probe.py ours 2>&1 | grep -v Warn | tee gpurun_out/r2i_ours.log | cut -c1-200
timeout 900 python tools/baseline_probe.py fa4 2>&1 | grep -v Warn | tee gpurun_out/r2i_fa4.log | cut -c1-200
timeout 600 python tools/bench_configs.py 2>&1 | grep -v Warn | tee gpurun_out/r2i_configs.txt | cut -c1-200

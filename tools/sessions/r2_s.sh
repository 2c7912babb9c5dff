#!/bin/bash
# round 2, GPU call S: new default forward (two-pass blocks, late row sum): full GPU suite, bench lines, ours vs FA4 (shim) on one lease
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2s_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2 | tee gpurun_out/r2s_smoke.log
timeout 600 python bench.py 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/r2s_bench.json | cut -c1-300
timeout 300 python bench.py --mode fwd 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/r2s_bench_fwd.json | cut -c1-300
timeout 300 python bench.py --causal 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/r2s_bench_causal.json | cut -c1-300
timeout 600 python tools/baseline_probe.py ours 2>&1 | grep -v "Warn\|warn" | tee gpurun_out/r2s_ours.log | cut -c1-200
timeout 900 python tools/baseline_probe.py fa4 2>&1 | grep -v "Warn\|warn" | tee gpurun_out/r2s_fa4.log | cut -c1-200

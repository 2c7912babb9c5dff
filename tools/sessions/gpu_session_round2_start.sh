#!/bin/bash
# first GPU call of the next round: full suite + long randomised stress incl. the head-dim 129..256 backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_pytest.log
timeout 400 python tools/stress_bwd.py 5 100 160,192,224,256,136 2>&1 | tail -12 | tee gpurun_out/r2_stress_hd256.log
timeout 300 python tools/stress_bwd.py 7 80 2>&1 | tail -6 | tee gpurun_out/r2_stress_bwd.log
timeout 200 python tools/probe_hd256_vs_fa2.py ours 2>&1 | tail -5
timeout 300 python bench.py 2>/dev/null | tail -1 | cut -c1-400

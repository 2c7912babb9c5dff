#!/bin/bash
# dropout bring-up + full regression
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropout_gpu.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/u_dropout.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_dropout_gpu.py 2>&1 | tail -8 | tee gpurun_out/u_full.log
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]" | tee gpurun_out/u_bwd_time.log
timeout 200 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/u_bench.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/f_pytest.log
timeout 600 python tools/bench_configs.py ours > gpurun_out/f_configs.jsonl 2> gpurun_out/f_configs.err; echo "configs exit $?"; cat gpurun_out/f_configs.jsonl; tail -3 gpurun_out/f_configs.err

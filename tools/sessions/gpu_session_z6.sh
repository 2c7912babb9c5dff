#!/bin/bash
mkdir -p gpurun_out
H=$PWD/flash_attention_b200
export FA_B200_BWD_PAIR=1
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py -q -m gpu 2>&1 | grep -E "^E |passed|failed|FAILED" | head -40

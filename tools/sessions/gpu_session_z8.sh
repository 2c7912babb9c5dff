#!/bin/bash
export FA_B200_BWD_PAIR=1
timeout 120 python tools/probe_det.py 2>&1 | tail -6
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py -q -m gpu 2>&1 | tail -4
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"

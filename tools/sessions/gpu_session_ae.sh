#!/bin/bash
for c in 0 4 8 16; do
  export FA_B200_SCHED_CHUNK=$c
  echo "== FA_B200_SCHED_CHUNK=$c"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]" | tail -2
done
export FA_B200_SCHED_CHUNK=8
timeout 300 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py -q -m gpu 2>&1 | tail -2

#!/bin/bash
for c in 0 4 8 16 32; do
  export FA_B200_SCHED_CHUNK_FWD=$c
  echo "== FA_B200_SCHED_CHUNK_FWD=$c"
  timeout 200 python tools/probe_fwd.py time 2>&1 | grep "time\]" | tail -2
done
export FA_B200_SCHED_CHUNK_FWD=8
timeout 600 python -m pytest tests/test_fwd_gpu.py tests/test_kvcache_gpu.py tests/test_fp8_gpu.py -q -m gpu 2>&1 | tail -2

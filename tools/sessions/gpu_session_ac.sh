#!/bin/bash
mkdir -p gpurun_out
for s in 0 "" 1; do
  if [ -z "$s" ]; then unset FA_B200_SCHED; else export FA_B200_SCHED=$s; fi
  echo "== FA_B200_SCHED='$s'"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
done
unset FA_B200_SCHED
timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_bwd_gpu.py tests/test_dropin_gpu.py tests/test_kvcache_gpu.py tests/test_fp8_gpu.py -q -m gpu 2>&1 | tail -4

#!/bin/bash
H=$PWD/flash_attention_b200
export FA_B200_LIB=$H/libfa_b200_guard.so
export FA_B200_FWD_TILES=2
for st in small masks gqa varlen split fp8; do
  echo "== stage $st (tiles=2, split softmax, guard)"
  timeout 150 python tools/probe_fwd.py $st 2>&1 | grep -E "\[|rror|timeout" | tail -12 | cut -c1-170
  rc=${PIPESTATUS[0]}
  if [ "$rc" != "0" ]; then echo "stage $st failed/hung rc=$rc; stop"; exit 0; fi
done
unset FA_B200_LIB FA_B200_FWD_TILES
for s in 1 0; do
  echo "== time FA_B200_FWD_SPLIT=$s"
  FA_B200_FWD_SPLIT=$s timeout 200 python tools/probe_fwd.py time 2>&1 | grep -E "time|TF" | tail -8 | cut -c1-200
done
FA_B200_FWD_TILES=2 timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_fp8_gpu.py tests/test_kvcache_gpu.py -q -m gpu 2>&1 | tail -3
FA_B200_LIB=$PWD/flash_attention_b200/libfa_b200_fts.so timeout 120 python tools/probe_fwd_ts.py 2>&1 | tail -4 | cut -c1-420

#!/bin/bash
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
export FA_B200_FWD_TILES=1
timeout 400 python -m pytest tests/test_fwd_gpu.py tests/test_kvcache_gpu.py tests/test_fp8_gpu.py tests/test_fa3_surface_gpu.py -x -q -m gpu > gpurun_out/p_pytest_t1.log 2>&1; echo "pytest(tiles=1, guard) exit $?"; tail -6 gpurun_out/p_pytest_t1.log
unset FA_B200_FWD_TILES
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest_auto.log 2>&1; echo "pytest(auto, guard) exit $?"; tail -4 gpurun_out/p_pytest_auto.log
unset FA_B200_LIB
for t in 1 2; do
  export FA_B200_FWD_TILES=$t
  echo "== FA_B200_FWD_TILES=$t"
  timeout 200 python tools/sweep.py ours 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l)
    if r['d']==128: print(r['causal'], r['seqlen'], 'fwd', r['fwd_tflops'])
"
done
unset FA_B200_FWD_TILES
timeout 300 python tools/bench_configs.py ours 2>/dev/null | grep -E "C5|C2|C4" | cut -c1-160

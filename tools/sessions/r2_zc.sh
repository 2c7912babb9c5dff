#!/bin/bash
# round 2, GPU call ZC: the whole GPU suite on the spin-guard build of the final sources (different timing of every mbarrier wait: the
# build that exposed the exchange-area race of the wide-S kernel), then this library's rows of the comparator table once more
mkdir -p gpurun_out
FA_B200_LIB=flash_attention_b200/variants/lib_wguard.so timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_dropin_versions_gpu.py 2>&1 | tail -5 | tee gpurun_out/r2zc_guard_suite.log
timeout 300 python tools/baseline_probe.py ours 2>&1 | grep -v "Warn\|warn" | tee gpurun_out/r2zc_ours.log | cut -c1-120

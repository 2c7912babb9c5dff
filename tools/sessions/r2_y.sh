#!/bin/bash
# round 2, GPU call Y: validation of the final build as the driver will run it: wide-kernel tests under the spin guard and with the
# one-pass option build, then the full GPU suite (incl. the drop-in tests on the 2.8.4 / FA3 interface files in baseline/_ref), smoke, bench
mkdir -p gpurun_out
V=flash_attention_b200/variants
FA_B200_LIB=$V/lib_wguard.so timeout 600 python -m pytest tests/test_fwd_wide_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/r2y_wide_guard.log
FA_B200_LIB=$V/lib_wop.so timeout 600 python -m pytest tests/test_fwd_wide_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/r2y_wide_onepass_build.log
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2 | tee gpurun_out/r2y_smoke.log
timeout 600 python bench.py 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/r2y_bench.json | cut -c1-300

#!/bin/bash
# round 2, GPU call C: compiled operator module end to end (full suite), drop-in on 3 reference interfaces, custom ops +
# torch.compile, host overhead, softmax rate probe
mkdir -p gpurun_out
timeout 200 ./experiments/softmax_rate_probe.bin 2>&1 | tee gpurun_out/r2c_softmax_rate_probe.txt
timeout 300 python tools/probe_overhead.py 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/r2c_overhead.log
timeout 900 python -m pytest tests/test_dropin_versions_gpu.py tests/test_custom_ops_gpu.py tests/test_dropin_gpu.py -q -x 2>&1 | tail -30 | tee gpurun_out/r2c_pytest_new.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2c_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

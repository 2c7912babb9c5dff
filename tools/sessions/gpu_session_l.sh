#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_pipeline_gpu.py tests/test_baseline_configs_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/l_bench.json')); print('value',d['value'],'e2e',d['e2e'],'roofline',d['roofline']['frac'],d['roofline']['traffic'],'clocks',d['clocks'])
PY
tail -3 gpurun_out/l_bench.err

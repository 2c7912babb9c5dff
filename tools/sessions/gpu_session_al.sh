#!/bin/bash
timeout 300 python -m pytest tests/test_host_pipeline_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e'])"

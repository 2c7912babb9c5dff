#!/bin/bash
for i in 1 2; do
for env in "" "FA_BENCH_NO_SAMPLER=1"; do
  for st in 10 20; do
    echo "== $env steps=$st"
    env $env timeout 200 python bench.py --steps $st --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f  bwd-alone %.1f  clocks %s' % (d['value'], d['roofline']['achieved'], d['clocks']))"
  done
done
done
timeout 200 python tools/bench_configs.py 2>&1 | grep headline | cut -c1-120

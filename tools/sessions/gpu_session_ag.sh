#!/bin/bash
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropout_gpu.py tests/test_dropin_gpu.py -q -m gpu 2>&1 | tail -3
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ag_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(open('gpurun_out/ag_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; col={h:i for i,h in enumerate(hdr)}
agg=defaultdict(list)
for r in rows[hi+1:]:
    if len(r)<len(hdr) or r[col['Metric Name']]!='gpu__time_duration.sum': continue
    agg[r[col['Kernel Name']][:50]].append(float(r[col['Metric Value']].replace(',','')))
for k,v in agg.items(): print(f"{k:52s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.1f} us")
PY

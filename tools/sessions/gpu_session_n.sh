#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep.py ours > gpurun_out/n_sweep_ours.jsonl 2> gpurun_out/n_sweep.err; echo "ours exit $?"; tail -2 gpurun_out/n_sweep.err
timeout 900 python tools/sweep.py fa2 > gpurun_out/n_sweep_fa2.jsonl 2>> gpurun_out/n_sweep.err; echo "fa2 exit $?"
python - <<'PY'
import json
o=[json.loads(l) for l in open('gpurun_out/n_sweep_ours.jsonl')]
f={(r['d'],r['causal'],r['seqlen']):r for r in map(json.loads, open('gpurun_out/n_sweep_fa2.jsonl'))}
for r in o:
    g=f.get((r['d'],r['causal'],r['seqlen']))
    print(r['d'],r['causal'],r['seqlen'],'fwd',r['fwd_tflops'],'fb',r['fwdbwd_tflops'],'| fa2', g and g['fwd_tflops'], g and g['fwdbwd_tflops'])
PY

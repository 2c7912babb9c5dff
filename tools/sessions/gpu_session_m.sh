#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/m_bench1.json 2> gpurun_out/m_bench1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/m_bench1.json')); print('N=1 value',d['value'],'e2e',d['e2e'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench2.json 2> gpurun_out/m_bench2.err; grep "^{" gpurun_out/m_bench2.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('N=2 value',d['value'],'n_gpus',d['n_gpus'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
"; tail -3 gpurun_out/m_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | grep "^{" | cut -c1-200

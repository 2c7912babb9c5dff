#!/bin/bash
# full validation of the default build (CTA-pair backward on) + evidence for profiles/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/ab_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 2>gpurun_out/ab_bench.err | tail -1 | tee gpurun_out/ab_bench.json | cut -c1-900
timeout 300 python bench.py --steps 10 --warmup 3 --causal --no-cpu-baseline 2>>gpurun_out/ab_bench.err | tail -1 | tee gpurun_out/ab_bench_causal.json | cut -c1-400
timeout 300 python tools/bench_configs.py 2>&1 | tail -8 | tee gpurun_out/ab_configs.txt | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_(fwd|bwd2)_kernel" -s 6 -c 2 -o gpurun_out/prof_r01_v5 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ab_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ab_ncu_list.log 2>&1; echo "ncu list exit $?"

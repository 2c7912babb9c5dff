#!/bin/bash
# tests + smoke + bench + profiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --causal --no-cpu-baseline > gpurun_out/bench_causal.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_causal.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_(fwd|bwd)_kernel" -s 6 -c 2 -o gpurun_out/prof_r01 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out

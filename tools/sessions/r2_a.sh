#!/bin/bash
# round 2, GPU call A: new parity tests (full-size vs FA2, fp8 golden, leftpad), same-box comparators (FA4 via shim, cuDNN,
# torch-SDPA-cuDNN, FA2, ours), ncu --set full of the kernels that had no capture in round 1, ex77 if built.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r2a_smi.txt
echo "== new tests first"
timeout 900 python -m pytest tests/test_fa2_fullsize_gpu.py tests/test_fp8_gpu.py tests/test_fwd_gpu.py::test_varlen_forward_leftpad_k tests/test_bwd_gpu.py tests/test_dropin_gpu.py -q -x -s 2>&1 | tail -25 | tee gpurun_out/r2a_pytest_new.log
echo "== comparators"
rm -f gpurun_out/r02_baseline.jsonl
timeout 300 python tools/baseline_probe.py ours 2>&1 | grep -v Warning | tail -20 | tee gpurun_out/r2a_probe_ours.log
timeout 900 python tools/baseline_probe.py fa4 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 | tee gpurun_out/r2a_probe_fa4.log
timeout 400 python tools/baseline_probe.py cudnn 2>&1 | grep -v Warning | tail -30 | tee gpurun_out/r2a_probe_cudnn.log
timeout 300 python tools/baseline_probe.py sdpa 2>&1 | grep -v Warning | tail -20 | tee gpurun_out/r2a_probe_sdpa.log
if [ -x baseline/_ref/ex77/77_blackwell_fmha_fp16 ]; then
  echo "== cutlass ex77"
  for m in no causal; do
    timeout 120 baseline/_ref/ex77/77_blackwell_fmha_fp16 --b=4 --h=16 --q=8192 --k=8192 --d=128 --mask=$m --iterations=10 2>&1 | tail -6
    timeout 120 baseline/_ref/ex77/77_blackwell_fmha_bwd_fp16 --b=4 --h=16 --q=8192 --k=8192 --d=128 --mask=$m --iterations=10 2>&1 | tail -6
  done | tee gpurun_out/r2a_ex77.log
fi
echo "== full suite"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2a_pytest_all.log
echo "== ncu"
# case:reps:skip:count  (skip/count are in launches that match the kernel filter)
for spec in fp8_c4:4:3:1 bf16_c4:4:3:1 decode_c5:4:6:2 bwd_hd256:3:8:4 hd64:3:8:4; do
  IFS=: read c reps skip cnt <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fa_fwd_kernel|fa_bwd|fa_combine' -s $skip -c $cnt -f -o gpurun_out/r2a_ncu_$c python tools/ncu_cases.py $c $reps > gpurun_out/r2a_ncu_$c.log 2>&1
  tail -2 gpurun_out/r2a_ncu_$c.log
done
timeout 300 python bench.py 2>gpurun_out/r2a_bench.err | tail -1 | tee gpurun_out/r2a_bench.json | cut -c1-1500

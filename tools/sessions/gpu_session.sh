#!/bin/bash
# run bring-up stages, each under its own timeout so a hang or trap cannot eat the GPU budget
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for st in "$@"; do
  case $st in
    fa2|fa4) timeout 600 python tools/baseline_probe.py $st > gpurun_out/baseline_$st.log 2>&1; echo "baseline $st exit $?" | tee -a gpurun_out/baseline_$st.log; tail -25 gpurun_out/baseline_$st.log;;
    bwd_*) timeout 200 python tools/probe_bwd.py ${st#bwd_} > gpurun_out/probe_$st.log 2>&1; echo "stage $st exit $?" | tee -a gpurun_out/probe_$st.log; tail -40 gpurun_out/probe_$st.log;;
    *) timeout 150 python tools/probe_fwd.py $st > gpurun_out/probe_$st.log 2>&1; echo "stage $st exit $?" | tee -a gpurun_out/probe_$st.log; tail -40 gpurun_out/probe_$st.log;;
  esac
done

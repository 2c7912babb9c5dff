#!/bin/bash
# v2 kernels: correctness with the spin-guard build, then timing of the poly-exp variants
mkdir -p gpurun_out
H=flash_attention_b200
export FA_B200_LIB=$PWD/$H/libfa_b200_guard.so
for st in small masks gqa varlen split paged fp8; do
  timeout 90 python tools/probe_fwd.py $st > gpurun_out/c_fwd_$st.log 2>&1; echo "fwd $st exit $?"; grep -E "FAIL|max_err|Error|error|timeout" gpurun_out/c_fwd_$st.log | head -20
done
for st in small masks gqa varlen; do
  timeout 100 python tools/probe_bwd.py $st > gpurun_out/c_bwd_$st.log 2>&1; echo "bwd $st exit $?"; grep -E "FAIL|Error|error|timeout|varlen" gpurun_out/c_bwd_$st.log | head -20
done
unset FA_B200_LIB
for v in "" _poly0 _poly1 _poly3; do
  export FA_B200_LIB=$PWD/$H/libfa_b200$v.so
  echo "== variant libfa_b200$v.so"
  timeout 200 python tools/probe_bwd.py time 2>&1 | grep time
done

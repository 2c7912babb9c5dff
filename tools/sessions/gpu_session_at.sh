#!/bin/bash
H=$PWD/flash_attention_b200
for v in "" _f8p1d2 _f8p2d2 _f8p3d2 _f8p4d2 _f8p2d3; do
  export FA_B200_LIB=$H/libfa_b200$v.so
  echo "== variant '$v'"
  timeout 200 python tools/probe_fwd.py fp8 2>&1 | grep -E "time C4|causal\]|window\]" | cut -c1-150
done

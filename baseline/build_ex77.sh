#!/bin/bash
# Builds the vendored CUTLASS example 77 (Blackwell FMHA forward / backward) from the reference tree where it lies
# (/root/reference/csrc/cutlass/examples/77_blackwell_fmha) into baseline/_ref/ex77/ -- a same-box tcgen05 comparator
# (BASELINE.md).  Binaries only; no reference source is copied.  baseline/_ref is git-ignored and travels to the GPU box.
set -e
CUT=/root/reference/csrc/cutlass
OUT=$(dirname "$0")/_ref/ex77
mkdir -p "$OUT"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr --expt-extended-lambda -DNDEBUG \
 -I$CUT/include -I$CUT/tools/util/include -I$CUT/examples/common -I$CUT/examples/77_blackwell_fmha -Xcompiler -fPIC -w"
for t in 77_blackwell_fmha 77_blackwell_fmha_bwd; do
  for prec in FP16; do
    nvcc $FLAGS -D$prec $CUT/examples/77_blackwell_fmha/$t.cu -o "$OUT/${t}_fp16" &
  done
done
wait
ls -la "$OUT"

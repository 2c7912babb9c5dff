#!/bin/bash
# round 2, GPU call Z: short / medium sequences: forward kernel-choice knobs (1-tile vs 2-tile crossover for causal, dynamic vs static deal)
mkdir -p gpurun_out
for cfg in "default:" "causal2tile:FA_B200_FWD_TILE_LIMIT_CAUSAL=0" "causal1tile16:FA_B200_FWD_TILE_LIMIT_CAUSAL=16" "causal1tile32:FA_B200_FWD_TILE_LIMIT_CAUSAL=32" "tiles1:FA_B200_FWD_TILES=1" "tiles2:FA_B200_FWD_TILES=2" "static:FA_B200_SCHED_DYNAMIC=0" "dynamic:FA_B200_SCHED_DYNAMIC=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "== $tag"; env $envs timeout 200 python tools/probe_short.py 2>&1 | grep -v Warn
done | tee gpurun_out/r2z_short_knobs.log

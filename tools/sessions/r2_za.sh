#!/bin/bash
# round 2, GPU call ZA: short / medium sequences: backward scheduling knobs
mkdir -p gpurun_out
for cfg in "default:" "static:FA_B200_SCHED_DYNAMIC=0" "group1:FA_B200_SCHED_GROUP_BWD=1" "group4:FA_B200_SCHED_GROUP_BWD=4" "group16:FA_B200_SCHED_GROUP_BWD=16" "group64:FA_B200_SCHED_GROUP_BWD=64" "single-cta:FA_B200_BWD_PAIR=0" "default2:"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "$tag: $(env $envs timeout 200 python tools/probe_short_bwd.py 2>&1 | grep -v Warn | tail -1)"
done | tee gpurun_out/r2za_short_bwd_knobs.log

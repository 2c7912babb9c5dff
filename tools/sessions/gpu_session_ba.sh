#!/bin/bash
mkdir -p gpurun_out
for c in 0 16; do
FA_B200_SCHED_CHUNK_FWD=$c timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"fa_fwd_kernel" -s 4 -c 1 --csv python bench.py --causal --mode fwd --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -E "dram__|gpu__time|hit_rate" | cut -d, -f5,13- | sed "s/^/chunk=$c /"
done

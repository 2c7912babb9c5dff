#!/bin/bash
# round 2, GPU call X: randomised stress of the final build (forward: two-tile default and the wide-S kernel forced; backward), new seeds
mkdir -p gpurun_out
(timeout 400 python tools/stress_fwd.py 21 50 2>&1 | tail -3
 FA_B200_FWD_WIDE=1 timeout 400 python tools/stress_fwd.py 22 50 2>&1 | tail -3
 timeout 500 python tools/stress_bwd.py 23 40 2>&1 | tail -3) | grep -v Warn | tee gpurun_out/r2x_stress.log

#!/bin/bash
# round 2, GPU call V (2 GPUs): the bench contract under torchrun at N = 2 (replicas, NCCL only for the barrier)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | grep -v "Warn\|warn" | tail -2 | tee gpurun_out/r2v_bench_n2.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep -v "Warn\|warn" | tail -1 | tee gpurun_out/r2v_bench_ref_n2.json | cut -c1-400

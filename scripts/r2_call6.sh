#!/bin/bash
# round 2, GPU call 6 (8 GPUs): phase timeline of the 8-GPU step, BASELINE config 5 on 8 GPUs, 8-rank parity
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 170 $TR --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-breakdown --timeline > gpurun_out/c6_n8.json 2> gpurun_out/c6_n8.err
timeout 240 $TR --master-port 29503 bench.py --gpus 8 --batch 131072 --dim 1024 --steps 5 --warmup 3 --no-kernel-breakdown > gpurun_out/c6_cfg5_n8.json 2> gpurun_out/c6_cfg5_n8.err
timeout 170 $TR --master-port 29504 tests/gpu_dist_check.py > gpurun_out/c6_dist_check.log 2>&1
tail -c 1500 gpurun_out/c6_n8.json; tail -5 gpurun_out/c6_n8.err; tail -c 1500 gpurun_out/c6_cfg5_n8.json; tail -5 gpurun_out/c6_cfg5_n8.err; tail -8 gpurun_out/c6_dist_check.log

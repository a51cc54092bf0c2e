#!/bin/bash
# round 2, GPU call 2 (8 GPUs): baseline of the round with a phase timeline, BASELINE config 5 on 8 GPUs, multi-rank parity
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 170 $TR --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-breakdown --timeline > gpurun_out/b2_n8.json 2> gpurun_out/b2_n8.err
timeout 170 $TR --master-port 29502 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-breakdown > gpurun_out/b2_n8_plain.json 2> gpurun_out/b2_n8_plain.err
timeout 240 $TR --master-port 29503 bench.py --gpus 8 --batch 131072 --dim 1024 --steps 5 --warmup 3 --no-kernel-breakdown > gpurun_out/b2_cfg5_n8.json 2> gpurun_out/b2_cfg5_n8.err
timeout 170 $TR --master-port 29504 tests/gpu_dist_check.py > gpurun_out/b2_dist_check.log 2>&1
tail -c 600 gpurun_out/b2_n8.json; tail -5 gpurun_out/b2_n8.err; tail -c 600 gpurun_out/b2_cfg5_n8.json; tail -5 gpurun_out/b2_cfg5_n8.err; tail -8 gpurun_out/b2_dist_check.log

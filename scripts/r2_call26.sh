#!/bin/bash
# round 2, GPU call 26 (1 GPU): ncu launch lists of the FINAL tree's default step and cfg-2 step (each after the same
# command has exited 0 without ncu)
mkdir -p gpurun_out
F="--steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-kernel-breakdown --no-parity"
python bench.py $F > gpurun_out/c26_plain_n1.json 2> gpurun_out/c26_plain_n1.err && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c26_launches_n1.csv \
    python bench.py $F > gpurun_out/c26_ncu_n1.log 2>&1
python bench.py --batch 4096 $F > gpurun_out/c26_plain_cfg2.json 2> gpurun_out/c26_plain_cfg2.err && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c26_launches_cfg2.csv \
    python bench.py --batch 4096 $F > gpurun_out/c26_ncu_cfg2.log 2>&1
wc -l gpurun_out/c26_launches_*.csv

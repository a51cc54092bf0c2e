#!/bin/bash
# round 2, GPU call 10 (1 GPU): GPU suite, ncu launch lists of the cfg-2 step and of the default step (clean: no parity / eager blocks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest.log 2>&1
tail -3 gpurun_out/c10_pytest.log
F="--no-cpu-baseline --no-gpu-eager --no-kernel-breakdown --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c10_launches_cfg2.csv \
    python bench.py --batch 4096 --steps 2 --warmup 3 $F > gpurun_out/c10_ncu_cfg2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c10_launches_n1.csv \
    python bench.py --steps 2 --warmup 3 $F > gpurun_out/c10_ncu_n1.log 2>&1
tail -n 2 gpurun_out/c10_ncu_cfg2.log gpurun_out/c10_ncu_n1.log

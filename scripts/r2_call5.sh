#!/bin/bash
# round 2, GPU call 5 (1 GPU): state of the tree after the re-entry -- full GPU suite, default bench line, cfg 2 / cfg 4
# lines, ncu launch list of the default step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1
tail -15 gpurun_out/c5_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err
tail -c 2500 gpurun_out/c5_bench_n1.json; tail -5 gpurun_out/c5_bench_n1.err
timeout 200 python bench.py --batch 4096 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/c5_bench_cfg2.json 2> gpurun_out/c5_bench_cfg2.err
tail -c 1500 gpurun_out/c5_bench_cfg2.json; tail -3 gpurun_out/c5_bench_cfg2.err
timeout 200 python bench.py --workload zeroshot --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c5_bench_zs.json 2> gpurun_out/c5_bench_zs.err
tail -c 800 gpurun_out/c5_bench_zs.json; tail -3 gpurun_out/c5_bench_zs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c5_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/c5_ncu_launches.log 2>&1
tail -3 gpurun_out/c5_ncu_launches.log

#!/bin/bash
# ncu evidence for the benchmark step (B200_PROFILING.md recipe): run under gpurun on ONE GPU, after the plain bench has
# exited 0.  Writes gpurun_out/{bench.json, launches.csv, top.ncu-rep}; summarise into profiles/ afterwards
# (ncu -i gpurun_out/top.ncu-rep --page raw --csv | grep -E 'dram__bytes|sm__pipe_tensor_cycles_active|gpu__time_duration').
#   gpurun --timeout 900 -- 'bash scripts/profile_step.sh'
set -e
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/ncu_launches.log 2>&1
PROBE_ONLY=stored PROBE_REPS=3 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_fused -s 1 -c 1 \
    -o gpurun_out/top python tests/gpu_stored_e_probe.py time > gpurun_out/ncu_top.log 2>&1
PROBE_ONLY=stored PROBE_REPS=3 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 \
    -o gpurun_out/fwd python tests/gpu_stored_e_probe.py time > gpurun_out/ncu_fwd.log 2>&1
ls -la gpurun_out

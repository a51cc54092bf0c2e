#!/bin/bash
# round 2, GPU call 15 (1 GPU): ncu --set full of the dominant kernel (fused backward) at the N=1 and the 8-GPU shard shape,
# and of the forward; each after the same command has exited 0 without ncu
mkdir -p gpurun_out
python tests/gpu_fused_once.py 32768 32768 2 > gpurun_out/c15_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_fused -s 1 -c 1 -o gpurun_out/c15_bwd_n1 -f \
    python tests/gpu_fused_once.py 32768 32768 2 > gpurun_out/c15_ncu_bwd_n1.log 2>&1
python tests/gpu_fused_once.py 4096 32768 2 >> gpurun_out/c15_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_fused -s 1 -c 1 -o gpurun_out/c15_bwd_shard -f \
    python tests/gpu_fused_once.py 4096 32768 2 > gpurun_out/c15_ncu_bwd_shard.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 1 -o gpurun_out/c15_fwd_n1 -f \
    python tests/gpu_fused_once.py 32768 32768 1 > gpurun_out/c15_ncu_fwd_n1.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/c15_plain.log gpurun_out/c15_ncu_bwd_n1.log

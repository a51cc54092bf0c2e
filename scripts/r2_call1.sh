#!/bin/bash
# round 2, GPU call 1: validate + time the kernel variants written after the round-1 GPU budget was spent
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a1_smi.log 2>&1
timeout 250 python tests/gpu_stored_e_probe.py variants > gpurun_out/a1_variants.log 2>&1
timeout 100 python tests/gpu_fused_probe.py 4096 32768 quick > gpurun_out/a1_shard_default.log 2>&1
MMG_FUSED_PANEL=1 timeout 100 python tests/gpu_fused_probe.py 4096 32768 quick > gpurun_out/a1_shard_panel.log 2>&1
timeout 120 python tests/gpu_rank_probe.py > gpurun_out/a1_rank.log 2>&1
tail -30 gpurun_out/a1_variants.log; cat gpurun_out/a1_shard_default.log gpurun_out/a1_shard_panel.log gpurun_out/a1_rank.log

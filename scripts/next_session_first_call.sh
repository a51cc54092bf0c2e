#!/bin/bash
# First GPU call of the next measurement session: validate + time the kernel variants that were written after the
# round-1 GPU budget was spent.  ~40 s of box time.
#   gpurun --timeout 300 -- 'bash scripts/next_session_first_call.sh'
mkdir -p gpurun_out
timeout 200 python tests/gpu_stored_e_probe.py variants 2>&1 | tee gpurun_out/variants.log | tail -40
# Then, on 2 GPUs (gpurun --gpus 2): the sharded loss with stored-E (opt-in) must still match the 1-GPU result, and time it:
#   MMGCLIP_B200_STORE_E_DIST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       tests/gpu_dist_check.py
#   MMGCLIP_B200_STORE_E_DIST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-kernel-breakdown

#!/bin/bash
# First GPU call of the next measurement session: validate + time the kernel variants that were written after the
# round-1 GPU budget was spent.  ~40 s of box time.
#   gpurun --timeout 300 -- 'bash scripts/next_session_first_call.sh'
mkdir -p gpurun_out
timeout 200 python tests/gpu_stored_e_probe.py variants 2>&1 | tee gpurun_out/variants.log | tail -40

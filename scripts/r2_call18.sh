#!/bin/bash
# round 2, GPU call 18 (2 GPUs): gather policy at the scaling-run shape (global batch 32768 on 2 ranks): push vs NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F="--no-kernel-breakdown --no-cpu-baseline --no-gpu-eager --no-parity"
MMGCLIP_B200_PUSH_GATHER=1 timeout 200 $TR --master-port 29501 bench.py --gpus 2 --steps 30 --warmup 5 $F > gpurun_out/c18_n2_push.json 2> gpurun_out/c18_n2_push.err
MMGCLIP_B200_PUSH_GATHER=0 timeout 200 $TR --master-port 29502 bench.py --gpus 2 --steps 30 --warmup 5 $F > gpurun_out/c18_n2_nccl.json 2> gpurun_out/c18_n2_nccl.err
MMGCLIP_B200_PUSH_GATHER=1 timeout 200 $TR --master-port 29503 bench.py --gpus 2 --steps 30 --warmup 5 $F > gpurun_out/c18_n2_push2.json 2> gpurun_out/c18_n2_push2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c18_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d["config"].get("gather", "")[:25])
    except Exception as e: print(f, "ERR", e)
PY

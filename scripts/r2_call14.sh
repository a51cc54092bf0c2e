#!/bin/bash
# round 2, GPU call 14 (2 GPUs): lighter push kernel (4 CTAs per destination) -- parity + A/B at 4096 rows per rank
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F="--no-kernel-breakdown --no-cpu-baseline --no-gpu-eager"
timeout 170 $TR --master-port 29504 tests/gpu_dist_check.py > gpurun_out/c14_dist_check.log 2>&1
tail -n 3 gpurun_out/c14_dist_check.log
timeout 200 $TR --master-port 29503 bench.py --gpus 2 --batch 8192 --steps 100 --warmup 5 $F > gpurun_out/c14_n2_b8192.json 2> gpurun_out/c14_n2_b8192.err
MMGCLIP_B200_PUSH_GATHER=0 timeout 200 $TR --master-port 29505 bench.py --gpus 2 --batch 8192 --steps 100 --warmup 5 $F --no-parity > gpurun_out/c14_n2_b8192_nccl.json 2> gpurun_out/c14_n2_b8192_nccl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c14_n2*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], (d.get("parity") or {}).get("ok"), d["config"].get("gather")[:30])
    except Exception as e: print(f, "ERR", e)
PY

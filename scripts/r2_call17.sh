#!/bin/bash
# round 2, GPU call 17 (8 GPUs): push all-gather vs NCCL all-gather at 8 ranks, timeline, BASELINE config 5, 8-rank parity
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
F="--no-kernel-breakdown --no-cpu-baseline --no-gpu-eager"
timeout 170 $TR --master-port 29501 bench.py --gpus 8 --steps 30 --warmup 5 $F --timeline > gpurun_out/c17_n8.json 2> gpurun_out/c17_n8.err
cp gpurun_out/timeline_n8.json gpurun_out/c17_timeline_n8.json
MMGCLIP_B200_PUSH_GATHER=0 timeout 170 $TR --master-port 29502 bench.py --gpus 8 --steps 30 --warmup 5 $F --no-parity > gpurun_out/c17_n8_nccl_gather.json 2> gpurun_out/c17_n8_nccl_gather.err
timeout 240 $TR --master-port 29503 bench.py --gpus 8 --batch 131072 --dim 1024 --steps 5 --warmup 3 $F > gpurun_out/c17_cfg5_n8.json 2> gpurun_out/c17_cfg5_n8.err
timeout 170 $TR --master-port 29504 tests/gpu_dist_check.py > gpurun_out/c17_dist_check.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c17_*.json")):
    if "timeline" in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], (d.get("parity") or {}).get("ok"), d["config"].get("gather", "")[:25])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c17_n8.err gpurun_out/c17_dist_check.log

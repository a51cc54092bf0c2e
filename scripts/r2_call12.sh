#!/bin/bash
# round 2, GPU call 12 (8 GPUs): the 8-GPU step after rotation / single all-reduce / ring buffers / two-stream heads / PDL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
F="--no-kernel-breakdown --no-cpu-baseline --no-gpu-eager"
timeout 170 $TR --master-port 29501 bench.py --gpus 8 --steps 30 --warmup 5 $F --timeline > gpurun_out/c12_n8.json 2> gpurun_out/c12_n8.err
timeout 170 $TR --master-port 29502 bench.py --gpus 8 --steps 30 --warmup 5 $F --no-parity --no-head-overlap --timeline > gpurun_out/c12_n8_nohead.json 2> gpurun_out/c12_n8_nohead.err
cp gpurun_out/timeline_n8.json gpurun_out/c12_timeline_n8_nohead.json
timeout 170 $TR --master-port 29504 tests/gpu_dist_check.py > gpurun_out/c12_dist_check.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/c12_n8.json", "gpurun_out/c12_n8_nohead.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], (d.get("parity") or {}).get("ok"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 4 gpurun_out/c12_n8.err gpurun_out/c12_dist_check.log

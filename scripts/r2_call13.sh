#!/bin/bash
# round 2, GPU call 13 (2 GPUs): push all-gather -- 2-rank parity (push, peer=0, NCCL gather fallback), 2-GPU bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/c13_pytest.log 2>&1
tail -8 gpurun_out/c13_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F="--no-kernel-breakdown --no-cpu-baseline --no-gpu-eager"
timeout 200 $TR --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 $F --timeline > gpurun_out/c13_n2.json 2> gpurun_out/c13_n2.err
MMGCLIP_B200_PUSH_GATHER=0 timeout 200 $TR --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 $F --no-parity > gpurun_out/c13_n2_nccl.json 2> gpurun_out/c13_n2_nccl.err
timeout 200 $TR --master-port 29503 bench.py --gpus 2 --batch 8192 --steps 50 --warmup 5 $F --timeline > gpurun_out/c13_n2_b8192.json 2> gpurun_out/c13_n2_b8192.err
cp gpurun_out/timeline_n2.json gpurun_out/c13_timeline_n2_b8192.json
MMGCLIP_B200_PUSH_GATHER=0 timeout 200 $TR --master-port 29504 bench.py --gpus 2 --batch 8192 --steps 50 --warmup 5 $F --no-parity > gpurun_out/c13_n2_b8192_nccl.json 2> gpurun_out/c13_n2_b8192_nccl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c13_n2*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], (d.get("parity") or {}).get("ok"), d["config"].get("gather"))
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/c13_timeline_n2_b8192.json").read())
PY
tail -n 3 gpurun_out/c13_n2.err gpurun_out/c13_n2_b8192.err

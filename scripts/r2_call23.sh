#!/bin/bash
# round 2, GPU call 23 (2 GPUs): the fp16 operand mode through the row-sharded path (gather of fp16 shards, peer reduction)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MMGCLIP_B200_EMB_F16=1 timeout 170 $TR --master-port 29504 tests/gpu_dist_check.py > gpurun_out/c23_dist_check_f16.log 2>&1
tail -n 3 gpurun_out/c23_dist_check_f16.log
timeout 200 $TR --master-port 29503 bench.py --gpus 2 --batch 8192 --steps 50 --warmup 5 --no-kernel-breakdown --no-cpu-baseline --no-gpu-eager --emb-f16 > gpurun_out/c23_n2_b8192_f16.json 2> gpurun_out/c23_n2_b8192_f16.err
python - <<'PY'
import json
f = "gpurun_out/c23_n2_b8192_f16.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["dtype"], d["ms_per_step"], d["value"], d["parity"]["ok"], d["parity"]["loss_rel_err"], d["parity"]["dw_image"], d["parity"]["dw_text"], d["config"]["gather"][:20])
except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c23_n2_b8192_f16.err

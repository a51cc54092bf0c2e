#!/bin/bash
# round 2, GPU call 8 (2 GPUs): multi-rank parity after the single-all-reduce loss / ring buffers / rotated column order,
# 2-GPU bench with the phase timeline, head-overlap A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1
tail -8 gpurun_out/c8_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-breakdown --no-cpu-baseline --timeline > gpurun_out/c8_n2_tl.json 2> gpurun_out/c8_n2_tl.err
timeout 200 $TR --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-breakdown --no-cpu-baseline --no-parity --head-overlap > gpurun_out/c8_n2_ho.json 2> gpurun_out/c8_n2_ho.err
python - <<'PY'
import json
for f in ("gpurun_out/c8_n2_tl.json", "gpurun_out/c8_n2_ho.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("parity"))
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/timeline_n2.json").read())
PY
tail -n 3 gpurun_out/c8_n2_tl.err gpurun_out/c8_n2_ho.err

#!/bin/bash
# round 2, GPU call 4 (2 GPUs): multi-rank parity after the distributed refactor, phase timeline of the 2-GPU step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_round2.py tests/test_gpu_stored_e.py tests/test_gpu_zeroshot_model.py -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1
tail -8 gpurun_out/c4_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-breakdown --timeline > gpurun_out/c4_n2_tl.json 2> gpurun_out/c4_n2_tl.err
timeout 200 $TR --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-breakdown > gpurun_out/c4_n2.json 2> gpurun_out/c4_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/c4_n2_tl.json", "gpurun_out/c4_n2.json"):
    try:
        d = json.load(open(f)); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["parity"])
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/timeline_n2.json").read())
PY
tail -3 gpurun_out/c4_n2.err

#!/bin/bash
# round 2, GPU call 16 (1 GPU): suite + N=1 / cfg-2 lines after the cluster loss kernels and the hint-plumbing removal
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c16_pytest.log 2>&1
tail -4 gpurun_out/c16_pytest.log
F="--no-cpu-baseline --no-gpu-eager --no-kernel-breakdown"
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 $F > gpurun_out/c16_cfg2.json 2> gpurun_out/c16_cfg2.err
timeout 120 python bench.py --steps 20 --warmup 5 $F > gpurun_out/c16_n1.json 2> gpurun_out/c16_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c16_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["parity"]["ok"], d["parity"]["loss_rel_err"], d["parity"]["dw_image"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c16_*.err

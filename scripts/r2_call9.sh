#!/bin/bash
# round 2, GPU call 9 (1 GPU): suite after the launch-count work (merged prep+diag, zero fill inside l2norm_bwd, loss parts),
# cfg 2 and N=1 lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1
tail -5 gpurun_out/c9_pytest.log
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-kernel-breakdown > gpurun_out/c9_cfg2.json 2> gpurun_out/c9_cfg2.err
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-kernel-breakdown > gpurun_out/c9_n1.json 2> gpurun_out/c9_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/c9_cfg2.json", "gpurun_out/c9_n1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], {k: v for k, v in d["parity"].items() if k != "checker"})
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c9_cfg2.err gpurun_out/c9_n1.err

#!/bin/bash
# round 2, GPU call 19 (1 GPU): fp16 embedding operands (MMG_PREC_F16) -- suite, smoke, cfg-2 and N=1 lines with the parity block
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c19_pytest.log 2>&1
tail -25 gpurun_out/c19_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c19_smoke.log 2>&1; tail -4 gpurun_out/c19_smoke.log
F="--no-cpu-baseline --no-gpu-eager --no-kernel-breakdown"
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 $F > gpurun_out/c19_cfg2.json 2> gpurun_out/c19_cfg2.err
timeout 120 python bench.py --steps 20 --warmup 5 $F > gpurun_out/c19_n1.json 2> gpurun_out/c19_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c19_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["parity"]["ok"], d["parity"]["loss_rel_err"], d["parity"]["dw_image"], d["parity"]["dw_text"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c19_*.err

#!/bin/bash
# round 2, GPU call 20 (1 GPU): all-fp16 InfoNCE mode (opt-in) -- suite (default bf16 path + the fp16 tests), cfg-2 / N=1 lines
# in both operand formats
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c20_pytest.log 2>&1
tail -12 gpurun_out/c20_pytest.log
F="--no-cpu-baseline --no-gpu-eager --no-kernel-breakdown"
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 $F --emb-f16 > gpurun_out/c20_cfg2_f16.json 2> gpurun_out/c20_cfg2_f16.err
timeout 120 python bench.py --steps 20 --warmup 5 $F --emb-f16 > gpurun_out/c20_n1_f16.json 2> gpurun_out/c20_n1_f16.err
timeout 120 python bench.py --steps 20 --warmup 5 $F > gpurun_out/c20_n1_bf16.json 2> gpurun_out/c20_n1_bf16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c20_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["dtype"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d["parity"]["ok"], d["parity"]["loss_rel_err"], d["parity"]["dw_image"], d["parity"]["dw_text"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c20_*.err

#!/bin/bash
# round 2, GPU call 21 (1 GPU): fp16-mode tests after the bar fix; operand-format A/B of the default step, alternating order
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/c21_pytest.log 2>&1
tail -6 gpurun_out/c21_pytest.log
F="--no-cpu-baseline --no-gpu-eager --no-kernel-breakdown --no-parity"
for i in 1 2; do
timeout 120 python bench.py --steps 30 --warmup 5 $F --emb-bf16 > gpurun_out/c21_n1_bf16_$i.json 2> gpurun_out/c21_n1_bf16_$i.err
timeout 120 python bench.py --steps 30 --warmup 5 $F --emb-f16 > gpurun_out/c21_n1_f16_$i.json 2> gpurun_out/c21_n1_f16_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c21_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["dtype"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["roofline"]["dominant_kernel_live"]["ms_per_launch"])
    except Exception as e: print(f, "ERR", e)
PY

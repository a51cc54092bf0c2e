#!/bin/bash
# round 2, GPU call 22 (1 GPU): the tree as committed -- suite, smoke(), default bench line (with CPU baseline), cfg-2 line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c22_pytest.log 2>&1
tail -6 gpurun_out/c22_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c22_smoke.log 2>&1; tail -4 gpurun_out/c22_smoke.log
timeout 300 python bench.py > gpurun_out/c22_bench_n1.json 2> gpurun_out/c22_bench_n1.err
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/c22_bench_cfg2.json 2> gpurun_out/c22_bench_cfg2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c22_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["dtype"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["parity"]["ok"], d["parity"]["dw_image"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/c22_*.err

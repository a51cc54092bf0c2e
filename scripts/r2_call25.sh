#!/bin/bash
# round 2, GPU call 25 (1 GPU): final tree -- suite, smoke(), default bench line, cfg-2 line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c25_pytest.log 2>&1
tail -4 gpurun_out/c25_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c25_smoke.log 2>&1; tail -2 gpurun_out/c25_smoke.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c25_bench_n1.json 2> gpurun_out/c25_bench_n1.err
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager > gpurun_out/c25_bench_cfg2.json 2> gpurun_out/c25_bench_cfg2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c25_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["dtype"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["parity"]["ok"], d["parity"]["dw_image"], d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 2 gpurun_out/c25_*.err

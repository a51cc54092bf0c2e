#!/bin/bash
# round 2, GPU call 7 (1 GPU): suite after the loss-part / ring / rotation changes, L2-hint + plan sweep of the fused backward,
# head-overlap A/B at the cfg-2 shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1
tail -5 gpurun_out/c7_pytest.log
timeout 200 python tests/gpu_hint_probe.py 32768 32768 > gpurun_out/c7_hint_32768.log 2>&1
timeout 200 python tests/gpu_hint_probe.py 4096 32768 > gpurun_out/c7_hint_4096x32768.log 2>&1
cat gpurun_out/c7_hint_32768.log gpurun_out/c7_hint_4096x32768.log
for ho in --no-head-overlap --head-overlap; do
timeout 120 python bench.py --batch 4096 --steps 100 --warmup 10 --no-cpu-baseline --no-parity --no-gpu-eager --no-kernel-breakdown $ho > gpurun_out/c7_cfg2$ho.json 2> gpurun_out/c7_cfg2$ho.err
python -c "import json,sys; d=json.loads(open('gpurun_out/c7_cfg2$ho.json').read().strip().splitlines()[-1]); print('$ho', d['ms_per_step'], d['e2e']['ms_per_step'])"
done

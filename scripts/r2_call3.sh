#!/bin/bash
# round 2, GPU call 3 (1 GPU): full GPU test suite after the refactor + the new round-2 tests, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1
tail -15 gpurun_out/c3_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err
tail -c 1500 gpurun_out/c3_bench_n1.json; tail -5 gpurun_out/c3_bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err
tail -c 600 gpurun_out/c3_bench_ref.json

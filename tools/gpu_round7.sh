#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu7.log
tail -6 gpurun_out/pytest_gpu7.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dff7.json 2> gpurun_out/bench_dff7.err
cat gpurun_out/bench_dff7.json; tail -3 gpurun_out/bench_dff7.err
ACCEL_BRANCHES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dff7_nobr.json 2> gpurun_out/bench_dff7_nobr.err
ACCEL_TAIL_STAGED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dff7_notail.json 2> gpurun_out/bench_dff7_notail.err
timeout 600 python tools/layer_times.py --version dff > gpurun_out/layer_times_dff7.txt 2>&1

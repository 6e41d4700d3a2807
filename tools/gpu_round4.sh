#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu4.log
tail -8 gpurun_out/pytest_gpu4.log
timeout 600 python tools/layer_times.py --version dff > gpurun_out/layer_times_dff4.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dff4.json 2> gpurun_out/bench_dff4.err
cat gpurun_out/bench_dff4.json; tail -3 gpurun_out/bench_dff4.err
ACCEL_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dff4_nopdl.json 2> gpurun_out/bench_dff4_nopdl.err
cat gpurun_out/bench_dff4_nopdl.json

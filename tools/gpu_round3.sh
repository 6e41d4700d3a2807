#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log
tail -8 gpurun_out/pytest_gpu3.log
( ACCEL_WARP_CB=4 timeout 120 python tools/bench_warp.py
  ACCEL_WARP_CB=8 timeout 120 python tools/bench_warp.py
  ACCEL_WARP_CB=4 ACCEL_WARP_THREADS=256 timeout 120 python tools/bench_warp.py ) > gpurun_out/bench_warp3.txt 2>&1
cat gpurun_out/bench_warp3.txt
timeout 600 python tools/layer_times.py --version dff > gpurun_out/layer_times_dff3.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dff3.json 2> gpurun_out/bench_dff3.err
cat gpurun_out/bench_dff3.json; tail -3 gpurun_out/bench_dff3.err

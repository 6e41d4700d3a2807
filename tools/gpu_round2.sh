#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_io.py tests/test_gpu_ops.py -m gpu -x -q -k "warp or io or preprocess or confusion or pipeline" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
( ACCEL_WARP_GATHER=1 timeout 120 python tools/bench_warp.py
  ACCEL_WARP_THREADS=256 timeout 120 python tools/bench_warp.py
  ACCEL_WARP_THREADS=512 timeout 120 python tools/bench_warp.py
  ACCEL_WARP_THREADS=512 ACCEL_WARP_TH=4 ACCEL_WARP_RMAX=9 timeout 120 python tools/bench_warp.py
  ACCEL_WARP_THREADS=256 ACCEL_WARP_TH=4 ACCEL_WARP_RMAX=9 timeout 120 python tools/bench_warp.py ) > gpurun_out/bench_warp2.txt 2>&1
cat gpurun_out/bench_warp2.txt
timeout 600 python tools/layer_times.py --version dff > gpurun_out/layer_times_dff.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_dff2.json 2> gpurun_out/bench_dff2.err
cat gpurun_out/bench_dff2.json; tail -3 gpurun_out/bench_dff2.err

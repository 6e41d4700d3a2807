#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu8.log
tail -6 gpurun_out/pytest_gpu8.log
for v in dff 18 50 101; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --version $v --multi-stream 2 > gpurun_out/bench_${v}_8.json 2> gpurun_out/bench_${v}_8.err
done
ACCEL_BRANCHES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --version 18 > gpurun_out/bench_18_8_nobr.json 2> gpurun_out/bench_18_8_nobr.err

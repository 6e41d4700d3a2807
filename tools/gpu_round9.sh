#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu9.log
tail -6 gpurun_out/pytest_gpu9.log
for v in dff 18 101; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --version $v > gpurun_out/bench_${v}_9.json 2> gpurun_out/bench_${v}_9.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lookahead > gpurun_out/bench_dff_9_nola.json 2> gpurun_out/bench_dff_9_nola.err
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dff_9_long.json 2> gpurun_out/bench_dff_9_long.err

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu5.log
tail -12 gpurun_out/pytest_gpu5.log
timeout 600 python tools/bench_layer.py --set res4_2c,res4_2a,res4_2b,res2_2c,res3_2c,res2_2b,res5_2c,fc6 --sweep debug 2> gpurun_out/layer_debug2.txt
grep -E "ACCEL_LAYER|FAILED" gpurun_out/layer_debug2.txt | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dff5.json 2> gpurun_out/bench_dff5.err
cat gpurun_out/bench_dff5.json; tail -3 gpurun_out/bench_dff5.err

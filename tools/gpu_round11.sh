#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu11.log
tail -6 gpurun_out/pytest_gpu11.log
timeout 600 python tools/layer_times.py --version dff > gpurun_out/layer_times_dff11.txt 2>&1
grep -E "key frame|cur frame|res2a_branch2c|res3b1_branch2c|res4b5_branch2|res5b_branch2c|im2col|conv1 |flow_conv1|maxpool|fc6" gpurun_out/layer_times_dff11.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dff_12.json 2> gpurun_out/bench_dff_12.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --version 101 > gpurun_out/bench_101_12.json 2> gpurun_out/bench_101_12.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --version 18 > gpurun_out/bench_18_12.json 2> gpurun_out/bench_18_12.err

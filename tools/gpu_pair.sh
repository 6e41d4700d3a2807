#!/bin/bash
set -x
mkdir -p gpurun_out
ACCEL_TC_PAIR=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv or deconv or deform" > gpurun_out/pytest_pair.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pair.log
tail -15 gpurun_out/pytest_pair.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 600 python tools/bench_layer.py --set res4_2c,res4_2a,res4_2b,res3_2c,res3_2b,res5_2a,res5_2c,fc6,flow_conv3 --sweep pair 2> gpurun_out/layer_pair.txt
grep -E "ACCEL_LAYER|FAILED" gpurun_out/layer_pair.txt | cut -c1-200

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "deform" > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu10.log
tail -6 gpurun_out/pytest_gpu10.log
timeout 600 python tools/layer_times.py --version dff 2>&1 | grep -E "im2col|key frame|cur frame" > gpurun_out/layer_times_dcn.txt
ACCEL_DCN_STAGED=0 timeout 600 python tools/layer_times.py --version dff 2>&1 | grep -E "im2col|key frame" >> gpurun_out/layer_times_dcn.txt
cat gpurun_out/layer_times_dcn.txt




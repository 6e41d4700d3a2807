#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graphs.py tests/test_gpu_linear_head.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_interval.py -x -q > gpurun_out/r02_t_warp.log 2>&1; tail -6 gpurun_out/r02_t_warp.log
echo "== bench_warp fused"; python tools/bench_warp.py 2>&1 | tail -7
echo "== bench_warp staged (old)"; ACCEL_WARP_FUSED=0 python tools/bench_warp.py 2>&1 | tail -7
python tools/layer_times.py --version dff --reps 3 2>&1 | grep -E "warp|==== "

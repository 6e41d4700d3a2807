#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "warp" 2>&1 | tail -5
timeout 600 python tools/bench_warp.py 2>&1 | tee gpurun_out/r02_bench_warp_fused.txt

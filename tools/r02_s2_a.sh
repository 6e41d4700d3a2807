#!/bin/bash
# session-2 call A: tcgen05 issue-rate probe + validation of the fused warp kernel's all-taps-in-range fast path
mkdir -p gpurun_out
tools/_build/mma_probe > gpurun_out/r02_mma_probe.txt 2>&1
cat gpurun_out/r02_mma_probe.txt
python -m pytest tests/test_gpu_ops.py -x -q -k "warp" 2>&1 | tail -3
python tools/bench_warp.py > gpurun_out/r02_bench_warp_fast.txt 2>&1
tail -5 gpurun_out/r02_bench_warp_fast.txt

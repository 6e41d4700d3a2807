#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu6.log
tail -6 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py --steps 10 --warmup 3 --multi-stream 2 > gpurun_out/bench_dff6.json 2> gpurun_out/bench_dff6.err
cat gpurun_out/bench_dff6.json; tail -3 gpurun_out/bench_dff6.err
timeout 600 python bench.py --steps 10 --warmup 3 --multi-stream 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dff6_ms3.json 2> gpurun_out/bench_dff6_ms3.err
timeout 600 python bench.py --steps 10 --warmup 3 --version 18 --no-cpu-baseline --multi-stream 2 > gpurun_out/bench_18_6.json 2> gpurun_out/bench_18_6.err
timeout 600 python bench.py --steps 10 --warmup 3 --version 101 --no-cpu-baseline --multi-stream 2 > gpurun_out/bench_101_6.json 2> gpurun_out/bench_101_6.err
for k in 1 2 5 10; do timeout 600 python bench.py --steps 10 --warmup 3 --version 50 --interval $k --no-cpu-baseline > gpurun_out/bench_50_k$k.json 2> gpurun_out/bench_50_k$k.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_kernel_staged -c 2 -f -o gpurun_out/warp_staged_full python tools/profile_step.py --version dff --intervals 1 --flags 2 > gpurun_out/ncu_warp_staged.log 2>&1
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json

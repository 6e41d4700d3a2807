#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list and full captures (outputs under gpurun_out/).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_dff.json 2> gpurun_out/bench_dff.err
timeout 600 python bench.py --steps 10 --warmup 3 --version 18 --no-cpu-baseline > gpurun_out/bench_18.json 2> gpurun_out/bench_18.err
timeout 600 python bench.py --steps 10 --warmup 3 --version 101 --no-cpu-baseline > gpurun_out/bench_101.json 2> gpurun_out/bench_101.err
timeout 300 python tools/bench_warp.py > gpurun_out/bench_warp.txt 2>&1
# launch list of one DFF key interval (second interval; graphs disabled so every node is a plain launch)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_dff.csv \
    python tools/profile_step.py --version dff --intervals 2 --flags 2 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_kernel -c 2 -f -o gpurun_out/warp_full \
    python tools/profile_step.py --version dff --intervals 1 --flags 2 > gpurun_out/ncu_warp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 6 -f -o gpurun_out/conv_full \
    python tools/profile_step.py --version dff --intervals 1 --flags 2 > gpurun_out/ncu_conv.log 2>&1
cat gpurun_out/bench_dff.json

#!/bin/bash
# round-2 profiler evidence: launch list of the bench command, full captures of the warp kernels and of res4 convs
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3800 --csv --log-file gpurun_out/r02_launches_bench101.csv \
  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc $?"; grep -c '^"' gpurun_out/r02_launches_bench101.csv
# warp kernels: fused (in the DFF cur plan) and the stand-alone staged kernel (accel_warp), rotating buffers, caches left alone
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:warp_kernel_fused -s 2 -c 4 -o gpurun_out/r02_warp_fused \
  python tools/profile_step.py --version dff --intervals 2 --flags 2 > gpurun_out/r02_ncu_warp_fused.log 2>&1; echo "fused rc $?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:warp_kernel_staged -s 12 -c 6 -o gpurun_out/r02_warp_staged \
  python tools/bench_warp.py > gpurun_out/r02_ncu_warp_staged.log 2>&1; echo "staged rc $?"
# res4 convs (2a 1x1 reduce, 2b 3x3, 2c 1x1 expand), single frame and five frames batched (stacked along H)
ACCEL_LAYER_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 12 -o gpurun_out/r02_conv_res4 \
  python tools/bench_layer.py --sweep one --reps 1 --set res4_2a,res4_2b,res4_2c,res4_2a_x5,res4_2b_x5,res4_2c_x5 > gpurun_out/r02_ncu_conv.log 2>&1; echo "conv rc $?"
ls -la gpurun_out/*.ncu-rep

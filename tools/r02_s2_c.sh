#!/bin/bash
SET=${1:-flow_conv3,res4_2a,res4_2b,res5_2a,res3_2b,res2_2b,res3_2a,res2_2b_x5,res4_2c_x5,res2_2c_x5}
for rep in 1 2; do
for lib in accel_b200/libaccel_b200_prev.so accel_b200/libaccel_b200.so; do
  echo "== $lib"
  ACCEL_B200_LIB=$PWD/$lib timeout 600 python tools/bench_layer.py --sweep one --set $SET 2>&1
done
done | tee gpurun_out/r02_layer_ab_single.txt

#!/bin/bash
mkdir -p gpurun_out
python tools/bench_layer.py --sweep r02 --set res2_2b,res3_2b,res4_2b,res5_off,flow_conv3_1 2> gpurun_out/r02_layer_slab2.txt; cat gpurun_out/r02_layer_slab2.txt

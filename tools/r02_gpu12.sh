#!/bin/bash
mkdir -p gpurun_out
python tools/bench_layer.py --sweep one --set res2_2a,res2_2a_x5,res3_2a,res3_2a_x5,res3_2b,res3_2b_x5,res4_2a,res4_2a_x5,res4_2b,res4_2b_x5,res4_2c,res4_2c_x5,res5_2a,res5_2a_x5,res5_2c,res5_2c_x5,fc6,fc6_x5,flow_conv4_1,flow_conv4_1_x4,flow_conv5_1,flow_conv5_1_x4,flow_conv6_1,flow_conv6_1_x4 2> gpurun_out/r02_layer_batch_potential.txt; cat gpurun_out/r02_layer_batch_potential.txt

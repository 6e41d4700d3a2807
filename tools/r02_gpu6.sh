#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "wide or conv_layer_matches or deconv" > gpurun_out/r02_t_wide.log 2>&1; tail -15 gpurun_out/r02_t_wide.log
python tools/bench_layer.py --sweep r02 --set res2_2b,res3_2b,res4_2a,res4_2b,res4_2c,res5_off,flow_conv3_1,fc6 2> gpurun_out/r02_layer_slab_chains.txt; cat gpurun_out/r02_layer_slab_chains.txt

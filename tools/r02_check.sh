#!/bin/bash
timeout 600 python tools/bench_layer.py --sweep mma --set res5_2a_x5,res4_2b_x5 2>&1 | tee gpurun_out/r02_layer_mma_only.txt

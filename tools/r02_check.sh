#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_linear_head.py tests/test_gpu_layers.py -x -q 2>&1 | tail -3

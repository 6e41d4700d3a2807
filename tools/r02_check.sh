#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_io.py tests/test_gpu_loader.py -x -q 2>&1 | tail -3

#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_io.py -x -q -k "whole_intervals" 2>&1 | tail -3
bash tools/r02_sanitize.sh

#!/bin/bash
mkdir -p gpurun_out
python tools/tc_rounding_probe.py > gpurun_out/r02_rounding_probe.txt 2>&1; cat gpurun_out/r02_rounding_probe.txt
python tools/debug_interval101.py 101 0 2>&1 | tail -4
ACCEL_BRANCHES=0 python tools/debug_interval101.py 101 0 2>&1 | tail -4
python tools/debug_interval101.py 101 2 2>&1 | tail -4
ACCEL_BRANCHES=0 python tools/debug_interval101.py 101 2 2>&1 | tail -4

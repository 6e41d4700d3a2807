#!/bin/bash
mkdir -p gpurun_out
python tools/tc_rounding_probe.py > gpurun_out/r02_rounding_probe.txt 2>&1; tail -22 gpurun_out/r02_rounding_probe.txt
python tools/debug_interval101.py 101 2 2>&1 | tail -6
ACCEL_IVL_MIN_SMS=148 python tools/debug_interval101.py 101 2 2>&1 | tail -6 | head -3

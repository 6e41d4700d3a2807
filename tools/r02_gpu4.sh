#!/bin/bash
mkdir -p gpurun_out
python tools/tc_rounding_probe.py > gpurun_out/r02_rounding_probe_chains.txt 2>&1; grep -A1 "k3 \|k1 " gpurun_out/r02_rounding_probe_chains.txt | grep -v "^--" | head -40
ACCEL_DEBUG_SUMS=1 python tools/debug_interval101.py 101 2 > gpurun_out/r02_dbg101.out 2> gpurun_out/r02_dbg101.err; grep "rep 0" gpurun_out/r02_dbg101.out
timeout 900 python -m pytest tests/test_gpu_fullsize_oracle.py -q -s > gpurun_out/r02_t2.log 2>&1; echo "exit $?" >> gpurun_out/r02_t2.log
grep -E "full-size parity|score volume max-abs|AssertionError|passed|failed|exit" gpurun_out/r02_t2.log | head -20
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/r02_t3.log 2>&1; tail -3 gpurun_out/r02_t3.log

#!/bin/bash
# compute-sanitizer over the operator-level GPU tests: memcheck on everything, racecheck on the
# kernels that synchronise through shared memory / mbarriers.  Summaries -> gpurun_out/, copied to profiles/.
#
# racecheck runs in three legs: (A) the subset with the fused warp kernel switched off (expected clean), (B) the fused warp
# kernel alone -- its producer/consumer ring is ordered by mbarriers only (cp.async.bulk writes vs generic reads), which
# racecheck reports as hazards -- and (C) tools/racecheck_probe.cu, a 60-line ring with the same protocol and a result
# check, to show what the tool says about that protocol in isolation.
mkdir -p gpurun_out tools/_build
export PATH=/usr/local/cuda/bin:$PATH
SUBSET="warp_staged_rows or tail_scores or conv_layer_matches or deconv_layer_matches or stem_matches or deformable_layer_matches or (interval_plan_matches and dff)"
if [ "$1" != "race" ]; then
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py tests/test_gpu_io.py tests/test_gpu_interval.py -x -q -k "not 2048-64-128 and not 64-128 and not full_size and not 1024-2048" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -6 gpurun_out/r02_sanitizer_memcheck.log
fi
L=gpurun_out/r02_sanitizer_racecheck.log
echo "=== leg A: ACCEL_WARP_FUSED=0" > $L
ACCEL_WARP_FUSED=0 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py tests/test_gpu_interval.py -x -q -k "$SUBSET" >> $L 2>&1
echo "racecheck leg A exit $?" >> $L
echo "=== leg B: fused warp kernel only" >> $L
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 4 \
  python -m pytest tests/test_gpu_interval.py -x -q -k "interval_plan_matches and dff" >> $L 2>&1
echo "racecheck leg B exit $?" >> $L
echo "=== leg C: tools/racecheck_probe.cu" >> $L
[ -x tools/_build/racecheck_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/_build/racecheck_probe tools/racecheck_probe.cu
tools/_build/racecheck_probe >> $L 2>&1
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 4 tools/_build/racecheck_probe >> $L 2>&1
echo "racecheck leg C exit $?" >> $L
grep -n "===\|exit\|SUMMARY\|passed\|failed\|ring_probe" $L

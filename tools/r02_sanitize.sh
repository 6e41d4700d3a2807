#!/bin/bash
# compute-sanitizer over the operator-level GPU tests: memcheck on everything, racecheck on the
# kernels that synchronise through shared memory / mbarriers.  Summaries -> gpurun_out/, copied to profiles/.
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py tests/test_gpu_io.py -x -q -k "not 2048-64-128 and not 64-128" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -6 gpurun_out/r02_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py -x -q -k "warp_staged_rows or tail_scores or conv_layer_matches or deconv_layer_matches or stem_matches or deformable_layer_matches" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -6 gpurun_out/r02_sanitizer_racecheck.log

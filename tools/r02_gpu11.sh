#!/bin/bash
mkdir -p gpurun_out
for cons in 256 512; do
echo "== fused warp CONS=$cons"
ACCEL_WARP_FUSED_CONS=$cons python tools/layer_times.py --version dff --reps 3 2>&1 | grep -E "warp"
ACCEL_WARP_FUSED_CONS=$cons timeout 300 python -m pytest tests/test_gpu_fullsize_oracle.py -q -k dff 2>&1 | tail -1
done
bash tools/r02_sanitize.sh

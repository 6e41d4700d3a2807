#!/bin/bash
# warp-per-(pixel, tap) deformable im2col (ACCEL_DCN_WARP=1): operator parity with the switch on, then a short A/B.
mkdir -p gpurun_out
ACCEL_DCN_WARP=1 timeout -k 5 40 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 30 -k "deformable" 2>&1 | tail -3
for m in 0 1; do
  ACCEL_DCN_WARP=$m timeout -k 5 40 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab6_$m.json 2>/dev/null
  python -c "import json; d=json.loads(open('gpurun_out/bench_ab6_$m.json').read()); print('dcn_warp=$m value %.1f key:backbone %s' % (d['value'], d['stage_ms_per_interval'].get('key:backbone')))"
done

#!/bin/bash
# session-2 call B: convergent MMA issuer / TMA producer -- correctness, per-layer A/B against the previous build, bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_interval.py tests/test_gpu_layers.py -x -q 2>&1 | tail -3
SET=res3_2b_x5,res4_2a_x5,res4_2b_x5,res5_2a_x5,res5_br1_x5,fc6,fc6_x5,flow_conv3_1,flow_conv4_1_x4,flow_conv6_1_x4,res5_off
for lib in accel_b200/libaccel_b200_prev.so accel_b200/libaccel_b200.so; do
  echo "== $lib"
  ACCEL_B200_LIB=$PWD/$lib timeout 600 python tools/bench_layer.py --sweep one --set $SET 2>&1
done | tee gpurun_out/r02_layer_issue_ab.txt
for i in 1 2; do
for lib in accel_b200/libaccel_b200_prev.so accel_b200/libaccel_b200.so; do
ACCEL_B200_LIB=$PWD/$lib timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/ab.json
python - "$lib" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
print(sys.argv[1],'value %.1f online %.1f lookahead %.1f'%(d['value'],d['online']['value'],d['lookahead']['value']), d['clocks']['sm_mhz'], d['parity']['key']['score_max_abs'] if 'parity' in d and d['parity'] else None)
PY
done
done 2>&1 | tee gpurun_out/r02_bench_issue_ab.txt

#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_interval.py -x -q 2>&1 | tail -3
timeout 600 python tools/bench_layer.py --sweep one --set res2_2c_x5,res2_2a_x5,res2_br1_x5,res3_2c_x5,res4_2a_x5,res4_2b_x5,res4_2c_x5,res5_2c_x5,res4_2c,res2_2b 2>&1 | tee gpurun_out/r02_layer_dma.txt
for i in 1 2; do
for lib in accel_b200/libaccel_b200_prev.so accel_b200/libaccel_b200.so; do
ACCEL_B200_LIB=$PWD/$lib timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/ab.json
python - "$lib" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
print(sys.argv[1],'value %.1f online %.1f lookahead %.1f'%(d['value'],d['online']['value'],d['lookahead']['value']), d['clocks']['sm_mhz'], d['parity']['key']['score_max_abs'] if 'parity' in d and d['parity'] else None)
PY
done
done

#!/bin/bash
# bench A/B of two builds of the library (same box, alternating): usage  bash tools/r02_s2_d.sh [bench args]
for i in 1 2; do
for lib in accel_b200/libaccel_b200_prev.so accel_b200/libaccel_b200.so; do
ACCEL_B200_LIB=$PWD/$lib timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e "$@" 2>/dev/null | tail -1 > gpurun_out/ab.json
python - "$lib" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
print(sys.argv[1],'value %.1f online %.1f lookahead %.1f'%(d['value'],d['online']['value'],d['lookahead']['value']), d['clocks']['sm_mhz'], d['parity']['key']['score_max_abs'] if 'parity' in d and d['parity'] else None)
PY
done
done 2>&1 | tee gpurun_out/r02_bench_ab.txt

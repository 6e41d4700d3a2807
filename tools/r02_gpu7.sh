#!/bin/bash
mkdir -p gpurun_out
for bo in 0 2; do
echo "== ASLAB_BO=$bo"
ACCEL_TC_ASLAB_BO=$bo timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "wide" 2>&1 | tail -3
done
export ACCEL_TC_ASLAB=0
echo "== chains single-wave only"
ACCEL_TC_CHAINS_MULTI=0 timeout 600 python -m pytest tests/test_gpu_fullsize_oracle.py -q -s 2>&1 | grep -E "full-size parity|passed|failed|Error" | cut -c1-400
ACCEL_TC_CHAINS_MULTI=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c_multi0.json 2>/dev/null
echo "== chains min 16"
ACCEL_TC_CHAINS_MIN=16 timeout 600 python -m pytest tests/test_gpu_fullsize_oracle.py -q -s 2>&1 | grep -E "full-size parity|passed|failed|Error" | cut -c1-400
ACCEL_TC_CHAINS_MIN=16 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c_min16.json 2>/dev/null
echo "== chains off"
ACCEL_TC_CHAINS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c_off.json 2>/dev/null
python - <<'PY'
import json
for n in ["multi0","min16","off"]:
    d=json.load(open('gpurun_out/r02_bench_c_%s.json'%n)); print(n, d['value'], d['ms_per_step'], d['online']['value'], d.get('lookahead',{}).get('value'))
PY

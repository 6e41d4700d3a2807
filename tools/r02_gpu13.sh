#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_interval.py -x -q -s > gpurun_out/r02_t_ivl.log 2>&1; grep -E "passed|failed|Error|error" gpurun_out/r02_t_ivl.log | head; grep -E "interval plan" gpurun_out/r02_t_ivl.log | cut -c1-160 | head -20
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench exit $?"; tail -3 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_d.json'))
for k in ['value','ms_per_step','launches_per_step','online','lookahead','accel18','dff','e2e','clocks','per_rank']:
    print(k, d.get(k))
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_full.log
tail -8 gpurun_out/r02_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_b.json'))
for k in ['value','ms_per_step','online','lookahead','accel18','dff','e2e','parity','clocks']:
    print(k, d.get(k))
PY

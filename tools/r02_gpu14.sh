#!/bin/bash
mkdir -p gpurun_out
python tools/interval_times.py --version 101 > gpurun_out/r02_interval_times_101.txt 2>&1; head -3 gpurun_out/r02_interval_times_101.txt
for ms in 8 24; do
echo "== MIN_SMS $ms"
ACCEL_IVL_MIN_SMS=$ms python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'])"
done

#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('fps', d['value'], d['ms_per_step'], 'online', d['online']['value'], 'look', d['lookahead']['value'])"; }
echo "== default (wide_kmax 16, multi_min 16)"
python tools/interval_parity.py 101 2>&1 | tail -1 | cut -c1-330
run
echo "== wide_kmax 40"
ACCEL_TC_WIDE_KMAX=40 python tools/interval_parity.py 101 2>&1 | tail -1 | cut -c1-330
ACCEL_TC_WIDE_KMAX=40 run

#!/bin/bash
# Commuted-L-head validation: full GPU suite (incl. tests/test_gpu_linear_head.py), then the bench with its
# `linear_head` extra object for DFF / Accel-18 / Accel-50.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_lin.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_lin.log
tail -15 gpurun_out/pytest_lin.log
for v in dff 18 50; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --version $v > gpurun_out/bench_lin_$v.json 2> gpurun_out/bench_lin_$v.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_lin_$v.json').read()); print('$v value %.1f e2e %.1f lin %s' % (d['value'], d['e2e']['value'], d.get('linear_head')))" || tail -5 gpurun_out/bench_lin_$v.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --linear-head > gpurun_out/bench_lin_main_dff.json 2> gpurun_out/bench_lin_main_dff.err
cat gpurun_out/bench_lin_main_dff.json | cut -c1-1500

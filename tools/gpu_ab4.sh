#!/bin/bash
# shared packed weights between plans: graph parity for every version, then DFF / Accel-101 benches (+ the strictly
# sequential online order, no key-frame lookahead).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graphs.py tests/test_golden.py tests/test_gpu_linear_head.py -m gpu -x -q > gpurun_out/pytest_ab4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ab4.log
tail -4 gpurun_out/pytest_ab4.log
run() {  # name version args...
  n=$1; v=$2; shift; shift
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --version $v "$@" > gpurun_out/bench_ab4_$n.json 2> gpurun_out/bench_ab4_$n.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_ab4_$n.json').read()); print('$n value %.1f ms/step %.3f lin %.1f' % (d['value'], d['ms_per_step'], (d.get('linear_head') or {}).get('value', 0)))" || tail -5 gpurun_out/bench_ab4_$n.err
}
run dff dff
run 101 101 --steps 10
run dff_online dff --no-lookahead
run 18 18

#!/bin/bash
# in-kernel split-K tail (ACCEL_TC_FUSED_SPLITK=1): operator + graph parity with the switch on, then same-box A/B.
mkdir -p gpurun_out
ACCEL_TC_FUSED_SPLITK=1 timeout -k 10 240 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graphs.py -m gpu -x -q --timeout 100 -k "not fold_on_and_off and not errors" > gpurun_out/pytest_ab5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ab5.log
tail -5 gpurun_out/pytest_ab5.log
run() {  # name version env args...
  n=$1; v=$2; e=$3; shift; shift; shift
  env $e timeout -k 10 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --version $v "$@" > gpurun_out/bench_ab5_$n.json 2> gpurun_out/bench_ab5_$n.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_ab5_$n.json').read()); print('$n value %.1f ms/step %.3f lin %.1f launches %d' % (d['value'], d['ms_per_step'], (d.get('linear_head') or {}).get('value', 0), d['launches_per_step']))" || tail -5 gpurun_out/bench_ab5_$n.err
}
run dff_sep dff ACCEL_TC_FUSED_SPLITK=0
run dff_fused dff ACCEL_TC_FUSED_SPLITK=1
run dff_online_fused dff ACCEL_TC_FUSED_SPLITK=1 --no-lookahead
run 18_fused 18 ACCEL_TC_FUSED_SPLITK=1

#!/bin/bash
# fc6 o feat_upsampling fold + band tail kernel: parity tests, then same-box A/B benches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graphs.py tests/test_golden.py tests/test_gpu_linear_head.py -m gpu -x -q > gpurun_out/pytest_fold_tail.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fold_tail.log
tail -12 gpurun_out/pytest_fold_tail.log
run() {  # name version env...
  n=$1; v=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --version $v > gpurun_out/bench_ft_$n.json 2> gpurun_out/bench_ft_$n.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_ft_$n.json').read()); print('$n value %.1f ms/step %.3f lin %.1f tail %s' % (d['value'], d['ms_per_step'], (d.get('linear_head') or {}).get('value', 0), d['stage_ms_per_interval'].get('cur:tail')))" || tail -5 gpurun_out/bench_ft_$n.err
}
run dff_old dff ACCEL_TAIL_BAND=0
run dff_new dff ACCEL_TAIL_BAND=1
run 18_old 18 ACCEL_TAIL_BAND=0 ACCEL_FOLD_FC6=0
run 18_fold 18 ACCEL_TAIL_BAND=0 ACCEL_FOLD_FC6=1
run 18_new 18 ACCEL_TAIL_BAND=1 ACCEL_FOLD_FC6=1
run 34_old 34 ACCEL_TAIL_BAND=0 ACCEL_FOLD_FC6=0
run 34_new 34 ACCEL_TAIL_BAND=1 ACCEL_FOLD_FC6=1

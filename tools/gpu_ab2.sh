#!/bin/bash
# band tail (rolled class loop) + flow-head SM budget: parity tests, then same-box A/B on the DFF bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graphs.py -m gpu -x -q > gpurun_out/pytest_ab2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ab2.log
tail -6 gpurun_out/pytest_ab2.log
run() {  # name version env...
  n=$1; v=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --version $v > gpurun_out/bench_ab2_$n.json 2> gpurun_out/bench_ab2_$n.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_ab2_$n.json').read()); s=d['stage_ms_per_interval']; print('$n value %.1f ms/step %.3f lin %.1f tail %s flownet %s' % (d['value'], d['ms_per_step'], (d.get('linear_head') or {}).get('value', 0), s.get('cur:tail'), s.get('cur:flownet')))" || tail -5 gpurun_out/bench_ab2_$n.err
}
run w4_t0 dff ACCEL_FLOWHEAD_WIDTH=4 ACCEL_TAIL_BAND=0
run w2_t0 dff ACCEL_FLOWHEAD_WIDTH=2 ACCEL_TAIL_BAND=0
run w2_t1 dff ACCEL_FLOWHEAD_WIDTH=2 ACCEL_TAIL_BAND=1
run w1_t1 dff ACCEL_FLOWHEAD_WIDTH=1 ACCEL_TAIL_BAND=1
run w4_t1 dff ACCEL_FLOWHEAD_WIDTH=4 ACCEL_TAIL_BAND=1
run 18_w2_t1 18 ACCEL_FLOWHEAD_WIDTH=2 ACCEL_TAIL_BAND=1

#!/bin/bash
# fused score-fusion tail: parity tests for the graph families that use it, benches, and one ncu --set full capture
# of tail_band_kernel (DFF: plain; Accel-18: with the in-kernel fusion).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphs.py tests/test_golden.py tests/test_gpu_ops.py -m gpu -x -q -k "not conv" > gpurun_out/pytest_ab3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ab3.log
tail -6 gpurun_out/pytest_ab3.log
run() {  # name version env...
  n=$1; v=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --version $v > gpurun_out/bench_ab3_$n.json 2> gpurun_out/bench_ab3_$n.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_ab3_$n.json').read()); s=d['stage_ms_per_interval']; print('$n value %.1f ms/step %.3f lin %.1f tail %s launches %d' % (d['value'], d['ms_per_step'], (d.get('linear_head') or {}).get('value', 0), s.get('cur:tail'), d['launches_per_step']))" || tail -5 gpurun_out/bench_ab3_$n.err
}
run 18_sep 18 ACCEL_TAIL_FUSE=0
run 18_fused 18 ACCEL_TAIL_FUSE=1
run 50_fused 50 ACCEL_TAIL_FUSE=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tail_band_kernel -s 2 -c 2 -o gpurun_out/tail_band_full -f \
    python tools/profile_step.py --version 18 --intervals 1 --flags 2 > gpurun_out/ncu_tail_band.log 2>&1
tail -2 gpurun_out/ncu_tail_band.log

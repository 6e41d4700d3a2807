#!/bin/bash
# A/B of one environment switch on the bench: usage  bash tools/gpu_ab.sh VERSION "ENV=VAL ..."
v=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphs.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for knobs in "" "$@"; do
  echo "== [$knobs]"
  env $knobs timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --version $v 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f  ms/step %.3f' % (d['value'], d['ms_per_step']))"
done

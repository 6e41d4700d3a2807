#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_golden.py tests/test_gpu_graphs.py -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu12.log
tail -4 gpurun_out/pytest_gpu12.log
for kmax in 10 8 18; do
echo "== KMAX $kmax"
ACCEL_TC_TMA_KMAX=$kmax timeout 600 python tools/layer_times.py --version dff 2>&1 | grep -E "key frame|cur frame"
ACCEL_TC_TMA_KMAX=$kmax timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], d['ms_per_step'])"
done

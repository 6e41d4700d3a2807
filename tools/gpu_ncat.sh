#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for n in 0 1; do
echo "== NCAT $n"
ACCEL_TC_NCAT=$n timeout 300 python tools/layer_times.py --version dff > gpurun_out/layer_times_ncat$n.txt 2>&1
grep -E "key frame|cur frame|res2a_branch2b|res3b1_branch2|res4b5_branch2|res5b_branch2a|flownet/conv2 |flownet/conv3 |conv3_1|conv4_1" gpurun_out/layer_times_ncat$n.txt
ACCEL_TC_NCAT=$n timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f  ms/step %.3f' % (d['value'], d['ms_per_step']))"
done

#!/bin/bash
set -x
mkdir -p gpurun_out
ACCEL_TC_TMA_OUT=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_tma.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tma.log
tail -8 gpurun_out/pytest_tma.log
for knobs in "" "ACCEL_TC_TMA_OUT=1" "ACCEL_TC_TMA_OUT=1 ACCEL_TC_BN=128" "ACCEL_TC_TMA_OUT=1 ACCEL_TC_BN=64" "ACCEL_TC_BN=128"; do
  echo "== $knobs"
  env $knobs ACCEL_LAYER_REPS=5 timeout 300 python - <<'PY' 2>&1 | grep ACCEL_LAYER | cut -c1-160
import sys, os
sys.path.insert(0, ".")
sys.argv = ["x"]
import torch
from accel_b200 import engine as E
sys.path.insert(0, "tools")
import bench_layer as B
for name in ["res4_2c", "res4_2a", "res4_2b", "res2_2c", "res3_2c", "res5_2c", "res2_2b", "fc6"]:
    cin, cout, h, w, k, s, p, d, res = B.LAYERS[name]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, cin, h, w, generator=g).cuda()
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    ho = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
    wo = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
    r = torch.randn(1, cout, ho, wo, generator=g).cuda() if res else None
    sys.stderr.write("%-10s " % name); sys.stderr.flush()
    E.conv_layer(x, wt, "conv", s, p, d, act=1, residual=r, engine=2)
PY
done

#!/usr/bin/env python
"""Debug aid: R101 of frame t as a key frame, GPU vs oracle: where do the features differ?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from accel_b200 import synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402
from oracle import nets, ops  # noqa: E402

H, W = 1024, 2048
params = synthetic.make_params("dff")
frames = synthetic.make_frames(10, H, W)
eng = Engine("dff", H, W, params=params)
dev = eng.torch_device
feat = torch.empty(eng.feat_shape, device=dev)
label = torch.empty(H, W, dtype=torch.uint8, device=dev)
for t in (2, 3):
    with torch.no_grad():
        rk = nets.key_forward(params, frames[t])
    eng.key_forward(frames[t].to(dev), feat, None, label)
    e = (feat.cpu() - rk["res5c_relu_output"]).abs()[0].max(dim=0).values.numpy()
    ys, xs = np.nonzero(e > 1e-2)
    print("frame %d as key: feat err max %.3e, feature px > 1e-2: %d" % (t, e.max(), len(ys)))
    if len(ys):
        print("   bbox y %d..%d x %d..%d; worst at %s" % (ys.min(), ys.max(), xs.min(), xs.max(), np.unravel_index(e.argmax(), e.shape)))

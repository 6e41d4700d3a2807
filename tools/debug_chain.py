#!/usr/bin/env python
"""Debug aid: where does the GPU chained loop leave the oracle's at 1024x2048?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from accel_b200 import synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402
from oracle import nets  # noqa: E402

version = sys.argv[1] if len(sys.argv) > 1 else "101"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 1024, 2048
params = synthetic.make_params(version)
frames = synthetic.make_frames(10, H, W)[:T]
eng = Engine(version, H, W, params=params)
dev = eng.torch_device
dfr = [f.to(dev) for f in frames]
feat = [torch.empty(eng.feat_shape, device=dev) for _ in range(2)]
score = torch.empty(1, 19, H, W, device=dev)
label = torch.empty(H, W, dtype=torch.uint8, device=dev)
with torch.no_grad():
    rk = nets.key_forward(params, frames[0])
eng.key_forward(dfr[0], feat[0], score, label)
print("frame 0: score err %.3e feat err %.3e" % ((score.cpu() - rk["croped_score_output"]).abs().max().item(),
                                                   (feat[0].cpu() - rk["res5c_relu_output"]).abs().max().item()))
rfeat = rk["res5c_relu_output"]
src = 0
for t in range(1, T):
    with torch.no_grad():
        rc = nets.cur_forward(params, version, frames[t], frames[t - 1], rfeat)
    eng.cur_forward(dfr[t], dfr[t - 1], feat[src], feat[src ^ 1], score, label)
    flow = eng.flownet(dfr[t], dfr[t - 1])
    e = (score.cpu() - rc[nets.output_key(version)]).abs()[0].max(dim=0).values.numpy()
    fe = (feat[src ^ 1].cpu() - rc["warping_feat_output"]).abs()[0].max(dim=0).values.numpy()
    fl = (flow.cpu() - rc["flow"]).abs().max().item()
    ys, xs = np.nonzero(e > 1e-2)
    print("frame %d: score err max %.3e, px > 1e-2: %d, > 1e-3: %d; warped feat err max %.3e (px > 1e-2: %d); flow err %.3e" % (
        t, e.max(), len(ys), int((e > 1e-3).sum()), fe.max(), int((fe > 1e-2).sum()), fl))
    if len(ys):
        print("   bbox of score err > 1e-2: y %d..%d x %d..%d" % (ys.min(), ys.max(), xs.min(), xs.max()))
        fy, fx = np.nonzero(fe > 1e-2)
        if len(fy):
            print("   warped-feature err > 1e-2 at feature px (y,x):", list(zip(fy.tolist(), fx.tolist()))[:12])
    # same cur frame fed the ORACLE's previous feature: isolates this frame's own error from the chain's
    eng.cur_forward(dfr[t], dfr[t - 1], rfeat.to(dev), feat[src], score, label)
    e2 = (score.cpu() - rc[nets.output_key(version)]).abs().max().item()
    print("   same frame on the oracle's previous feature: score err %.3e" % e2)
    eng.cur_forward(dfr[t], dfr[t - 1], feat[src ^ 1 if False else src], feat[src ^ 1], score, label) if False else None
    rfeat = rc["warping_feat_output"]
    # restore the GPU chain state (feat[src^1] holds the GPU's warped feature of frame t)
    src ^= 1

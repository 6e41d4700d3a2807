#!/usr/bin/env python
"""Per-layer parity of the R101-DCN key net at 1024x2048: every layer's internal output (accel_debug_fetch) against the
oracle's activation of the same name, on frame `t` of the benchmark's synthetic stream.  Prints the first layer whose
max-abs error leaves the noise floor and where.

    python tools/layer_parity.py [frame index ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from accel_b200 import synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402
from oracle import nets, ops  # noqa: E402


def main():
    H, W = 1024, 2048
    which = [int(a) for a in sys.argv[1:]] or [2]
    params = synthetic.make_params("dff")
    frames = synthetic.make_frames(10, H, W)
    eng = Engine("dff", H, W, params=params)
    dev = eng.torch_device
    feat = torch.empty(eng.feat_shape, device=dev)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    for t in which:
        with torch.no_grad():
            acts = nets.resnet_dcn_101_layers(params, frames[t])
        eng.key_forward(frames[t].to(dev), feat, None, label)
        torch.cuda.synchronize()
        print("==== frame %d" % t)
        reported = 0
        for name, ref in acts.items():
            got = eng.fetch_layer("key", name)[0].cpu()
            e = (got - ref[0]).abs()
            scale = ref.abs().max().item()
            emax = e.max().item()
            flag = ""
            if emax > 1e-3 * max(1.0, scale):
                pm = e.max(dim=0).values.numpy()
                ys, xs = np.nonzero(pm > 1e-3 * max(1.0, scale))
                c = int(e.reshape(e.shape[0], -1).max(dim=1).values.argmax())
                flag = "  <-- %d px off, bbox y %d..%d x %d..%d (map %dx%d), worst channel %d" % (
                    len(ys), ys.min(), ys.max(), xs.min(), xs.max(), e.shape[1], e.shape[2], c)
                reported += 1
            if flag or name in ("conv1", "maxpool") or name.endswith("c_branch2c") or name.endswith("a_branch2c"):
                print("%-28s |ref|max %9.3f  err max %.3e%s" % (name, scale, emax, flag))
            if reported >= 6:
                break


if __name__ == "__main__":
    main()

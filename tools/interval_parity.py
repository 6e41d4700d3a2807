#!/usr/bin/env python
"""Score error of every frame of one Accel-<v> key interval (whole-interval plan, 1024x2048) against the CPU oracle's
chained schedule -- the check behind the choice of ACCEL_TC_CHAINS_MULTI_MIN."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from accel_b200 import synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402
from oracle import ops  # noqa: E402
from oracle import schedule as oracle_schedule  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import critical_mask  # noqa: E402

version = sys.argv[1] if len(sys.argv) > 1 else "101"
I, H, W = 5, 1024, 2048
params = synthetic.make_params(version)
frames = synthetic.make_frames(2 * I, H, W)[:I]
eng = Engine(version, H, W, params=params, interval=I)
dev = eng.torch_device
dfr = [f.to(dev) for f in frames]
labels = torch.empty(I, H, W, dtype=torch.uint8, device=dev)
scores = [torch.empty(1, 19, H, W, device=dev) for _ in range(I)]
eng.interval_forward(dfr, labels, scores)
torch.cuda.synchronize()
ops.DCN_TRACE = []
with torch.no_grad():
    ref = oracle_schedule.run(params, version, frames, I, "chained", keep=("label", "score"))
trace, ops.DCN_TRACE = ops.DCN_TRACE, None
per = len(trace) // I if len(trace) % I == 0 else 0          # deformable layers per frame (Accel-101: 3 in every frame)
errs, errs_all, flips, excl = [], [], [], []
mask = None
for t in range(I):
    # a border-critical sample of frame t touches frame t and, through the chained feature, every later frame
    m = critical_mask(trace[t * per:(t + 1) * per] if per else trace, H, W)
    mask = m if mask is None else (mask | m)
    e = (scores[t].cpu() - ref[t]["score"]).abs()[0].max(dim=0).values.numpy()
    use = bool(mask.any()) and float(e.max()) >= 1e-3          # set the footprint aside only where a border tap flipped
    keep = ~mask if use else (mask | True)
    errs.append(float(e[keep].max()))
    errs_all.append(float(e.max()))
    excl.append(int((~keep).sum()))
    flips.append(int(((labels[t].cpu().numpy() != ref[t]["label"]) & keep).sum()))
print("env", {k: v for k, v in os.environ.items() if k.startswith("ACCEL_TC_CHAINS")}, version,
      "score max-abs per frame (outside the footprint of border-critical DCN samples)", ["%.2e" % e for e in errs],
      "incl. those px", ["%.2e" % e for e in errs_all], "excluded px", excl, "flipped labels", flips)

#!/usr/bin/env python
"""Per-op device times of the whole-interval plan (eager, CUDA events between ops, chains serialised)."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from accel_b200 import scheduler, synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--version", default="101")
ap.add_argument("--interval", type=int, default=5)
ap.add_argument("--top", type=int, default=0)
a = ap.parse_args()
H, W, I = 1024, 2048, a.interval
eng = Engine(a.version, H, W, params=synthetic.make_params(a.version), interval=I)
dev = eng.torch_device
fr = [f.to(dev) for f in synthetic.make_frames(I, H, W)]
labels = torch.empty(I, H, W, dtype=torch.uint8, device=dev)
st = scheduler.StreamState(eng)
best = None
for rep in range(4):
    eng.set_profiling(rep > 0)
    scheduler.segment_interval(eng, st, fr, labels)
    torch.cuda.synchronize()
    if rep > 0:
        t = eng.op_times()
        best = t if best is None else [(n, min(m, b[1]), fl) for (n, m, fl), b in zip(t, best)]
eng.set_profiling(False)
tot = sum(m for _, m, _ in best)
print("interval plan %s: %d ops, %.3f ms serialised" % (a.version, len(best), tot))
agg = collections.OrderedDict()
for n, m, fl in best:
    k = n.split("/")[0]
    agg[k] = agg.get(k, 0.0) + m
print({k: round(v, 3) for k, v in agg.items()})
for n, m, fl in best:
    if a.top == 0 or m > 0.05:
        rate = "%7.1f TF16/s" % (3 * fl / (m * 1e-3) / 1e12) if fl > 0 and m > 0 else ""
        print("%-52s %8.4f ms %8.2f GF %s" % (n[:52], m, fl / 1e9, rate))

#!/usr/bin/env python
"""Debug aid: Accel-101 whole-interval plan vs the frame-by-frame loop at 128x256, per frame."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from accel_b200 import scheduler, synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402

version = sys.argv[1] if len(sys.argv) > 1 else "101"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H, W, I = 128, 256, 3
eng = Engine(version, H, W, params=synthetic.make_params(version), interval=I, flags=flags)
dev = eng.torch_device
fr = [f.to(dev) for f in synthetic.make_frames(I, H, W, stream=2)]
labels = torch.empty(I, H, W, dtype=torch.uint8, device=dev)
scores = [torch.empty(1, 19, H, W, device=dev) for _ in range(I)]
st = scheduler.StreamState(eng)
st2 = scheduler.StreamState(eng)
lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
sc = [torch.empty(1, 19, H, W, device=dev) for _ in range(I)]
for t in range(I):
    scheduler.segment_frame(eng, st2, fr[t], I, "chained", lab, sc[t])
for rep in range(3):
    scheduler.segment_interval(eng, st, fr, labels, scores)
    torch.cuda.synchronize()
    print("env BRANCHES=%s flags=%d rep %d:" % (os.environ.get("ACCEL_BRANCHES"), flags, rep),
          ["%.3e" % (scores[t] - sc[t]).abs().max().item() for t in range(I)])
eng.set_profiling(True)
scheduler.segment_interval(eng, st, fr, labels, scores)
torch.cuda.synchronize()
ops = eng.op_times()
eng.set_profiling(False)
print("ops in the interval plan:", len(ops))
import collections
print(collections.Counter(n.split("/")[0] for n, _, _ in ops))
print([n for n, _, _ in ops if "corr" in n or "warp" in n or "res5c_branch2c" in n or "fc6" in n])

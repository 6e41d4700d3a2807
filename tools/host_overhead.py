#!/usr/bin/env python
"""Host-side cost of issuing one frame through scheduler.VideoPipeline / segment_frame (no GPU waits in the loop)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from accel_b200 import scheduler, synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402

H, W, I = 1024, 2048, 5
eng = Engine("dff", H, W, params=synthetic.make_params("dff"))
dev = eng.torch_device
u8 = synthetic.make_frames_u8(2 * I, H, W)
host = [f.contiguous().pin_memory() for f in u8]
frames = [synthetic.transform(f).to(dev) for f in u8]
labels = [torch.empty(H, W, dtype=torch.uint8).pin_memory() for _ in range(4)]
lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
state = scheduler.StreamState(eng)
for rep in range(3):
    for k in range(4 * I):
        scheduler.segment_frame(eng, state, frames[k % (2 * I)], I, "chained", lab)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20 * I
for k in range(n):
    scheduler.segment_frame(eng, state, frames[k % (2 * I)], I, "chained", lab)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("segment_frame: host issue %.1f us/frame, with GPU drain %.1f us/frame" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
pipe = scheduler.VideoPipeline(eng, I, "chained")
for rep in range(2):
    for k in range(4 * I):
        pipe.submit(host[k % (2 * I)], labels[k % 4], next_key_host=host[(k + I) % (2 * I)] if k % I == 0 else None)
pipe.sync()
t0 = time.perf_counter()
for k in range(n):
    pipe.submit(host[k % (2 * I)], labels[k % 4], next_key_host=host[(k + I) % (2 * I)] if k % I == 0 else None)
t1 = time.perf_counter()
pipe.sync()
t2 = time.perf_counter()
print("VideoPipeline.submit: host issue %.1f us/frame, with GPU drain %.1f us/frame" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))

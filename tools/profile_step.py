#!/usr/bin/env python
"""One warm-up key interval + one profiled key interval of the hot path, for ncu (see profiles/README.md).

    ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of warm-up> -c <N> --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --version dff
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from accel_b200 import scheduler, synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--version", default="dff")
    ap.add_argument("--interval", type=int, default=5)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--intervals", type=int, default=2)
    ap.add_argument("--flags", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = Engine(a.version, a.height, a.width, params=synthetic.make_params(a.version), device=0, flags=a.flags)
    frames = [f.to(dev) for f in synthetic.make_frames(a.interval, a.height, a.width)]
    label = torch.empty(a.height, a.width, dtype=torch.uint8, device=dev)
    state = scheduler.StreamState(eng)
    n = 0
    for s in range(a.intervals):
        for i in range(a.interval):
            scheduler.segment_frame(eng, state, frames[i], a.interval, "chained", label)
            n += eng.last_launch_count()
        torch.cuda.synchronize()
        print("interval %d done, launches so far %d" % (s, n), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-layer device times (CUDA events between consecutive launches) of one key and one cur frame.

    python tools/layer_times.py --version dff [--height 1024 --width 2048] [--reps 3]

Prints, per layer: best-of-reps milliseconds, reference-graph GFLOP, and the fp16-MMA rate it implies
(3 tensor-core passes per reference flop: fp16x3)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from accel_b200 import synthetic  # noqa: E402
from accel_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--version", default="dff")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--flags", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = Engine(a.version, a.height, a.width, params=synthetic.make_params(a.version), device=0, flags=a.flags)
    f = [x.to(dev) for x in synthetic.make_frames(2, a.height, a.width)]
    feat = [torch.empty(eng.feat_shape, device=dev) for _ in range(2)]
    label = torch.empty(a.height, a.width, dtype=torch.uint8, device=dev)
    for kind in ("key", "cur"):
        best = None
        for rep in range(a.reps + 1):
            eng.set_profiling(rep > 0)
            if kind == "key":
                eng.key_forward(f[0], feat[0], None, label)
            else:
                eng.cur_forward(f[1], f[0], feat[0], feat[1], None, label)
            torch.cuda.synchronize()
            if rep > 0:
                t = eng.op_times()
                best = t if best is None else [(n, min(m, b[1]), fl) for (n, m, fl), b in zip(t, best)]
        eng.set_profiling(False)
        total = sum(m for _, m, _ in best)
        print("==== %s frame: %d ops, %.3f ms (sum of per-op event intervals)" % (kind, len(best), total))
        for n, m, fl in best:
            rate = "%7.1f TF16/s" % (3 * fl / (m * 1e-3) / 1e12) if fl > 0 and m > 0 else ""
            print("%-52s %8.4f ms %8.2f GF %s" % (n[:52], m, fl / 1e9, rate))


if __name__ == "__main__":
    main()

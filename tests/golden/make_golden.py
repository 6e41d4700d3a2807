#!/usr/bin/env python
"""Generates tests/golden/accel_<version>_128x256.npz from the CPU oracle (CPU only, ~2 min).

The reference (SamvitJ/Accel) ships no golden vectors and MXNet cannot be imported here (SURVEY.md
section 8c).  This generator needs only the oracle; tests/golden/make_reference_wired.py is the one that wrote the
committed files: it produces the same arrays from the REFERENCE'S OWN symbol files (executed on oracle/mxstub.py),
after checking that they equal this oracle bit for bit.  These fixtures pin the ORACLE: seeded synthetic weights (accel_b200/synthetic.py,
seed 0) and a 3-frame synthetic clip (stream 0) through oracle/schedule.py with the reference's
chained loop (dff_deeplab/demo.py:228-250), interval 3.  Stored per frame: the uint8 label map, the
low-resolution quantities that determine it (flow, fp16-rounded full score volume is too large, so
the score volume is stored at every 8th pixel in fp32), and float64 checksums of the feature map.

    python tests/golden/make_golden.py            # rewrites the fixtures
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from accel_b200 import synthetic  # noqa: E402
from oracle import nets  # noqa: E402
from oracle import schedule as oracle_schedule  # noqa: E402

H, W, FRAMES, INTERVAL = 128, 256, 3, 3
SUB = 8          # score volume sub-sampling stride


def golden_for(version):
    params = synthetic.make_params(version)
    frames = synthetic.make_frames(FRAMES, H, W)
    with torch.no_grad():
        res = oracle_schedule.run(params, version, frames, INTERVAL, "chained", keep=("label", "score", "feat", "flow"))
    return pack(res)


def pack(res):
    """Per-frame results (dicts with 'label', 'score', 'feat', optional 'flow') -> the arrays of one fixture file."""
    out = {"height": H, "width": W, "interval": INTERVAL, "sub": SUB}
    for i, r in enumerate(res):
        score = r["score"]
        top2 = score.topk(2, dim=1).values
        out["label_%d" % i] = np.asarray(r["label"], dtype=np.uint8)
        out["margin_%d" % i] = (top2[:, 0] - top2[:, 1])[0].numpy().astype(np.float32)
        out["score_sub_%d" % i] = score[0, :, ::SUB, ::SUB].numpy().astype(np.float32)
        out["feat_sum_%d" % i] = np.float64(r["feat"].double().sum().item())
        out["feat_abs_sum_%d" % i] = np.float64(r["feat"].double().abs().sum().item())
        out["feat_sub_%d" % i] = r["feat"][0, ::64].numpy().astype(np.float32)        # 32 of 2048 channels
        if r.get("flow") is not None:
            out["flow_%d" % i] = r["flow"][0].numpy().astype(np.float32)
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for version in ("dff", "18", "34", "50", "101"):
        g = golden_for(version)
        path = os.path.join(HERE, "accel_%s_%dx%d.npz" % (version, H, W))
        np.savez_compressed(path, **g)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

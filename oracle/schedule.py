"""The reference's two keyframe loops restated over oracle.nets.  Test infrastructure.

chained   -- dff_deeplab/demo.py:165-250: `data_key` is the PREVIOUS frame and the feature carried
             to the next frame is the warped one (`warping_feat_output`).
unchained -- dff_deeplab/core/loader.py:259-303 + core/tester.py:246-256: `data_key` is the KEY
             frame image and `feat` stays the key frame's `res5c_relu_output` for the interval.
"""
from __future__ import annotations

from . import nets, ops


def key_frame_flags(num_frames, interval):
    """TestLoader's key_frame_flag for one video (loader.py:259-295): 0 first key frame, 1 later key
    frames, 2 non-key frames; key_frameid advances when cur - key == interval (:263-269)."""
    flags, key = [], 0
    for cur in range(num_frames):
        if cur - key == interval:
            key = cur
        flags.append((0 if key == 0 else 1) if key == cur else 2)
    return flags


def run(p, version, frames, interval, schedule="chained", keep=("label",)):
    """frames: list of (1,3,H,W) fp32 tensors (already `transform`ed).  Returns a list of dicts per
    frame with 'label' (uint8 HxW) and, when listed in `keep`, 'score', 'feat', 'flow'."""
    assert schedule in ("chained", "unchained")
    results = []
    feat = None
    key_img = None
    prev = None
    for idx, data in enumerate(frames):
        if prev is None:
            prev = data                                           # demo.py:178-179
        if idx % interval == 0:                                   # demo.py:235 / flag != 2
            out = nets.key_forward(p, data)
            feat = out["res5c_relu_output"]
            score = out["croped_score_output"]
            key_img = data
            flow = None
        else:
            data_key = prev if schedule == "chained" else key_img
            out = nets.cur_forward(p, version, data, data_key, feat)
            score = out[nets.output_key(version)]
            flow = out["flow"]
            if schedule == "chained":
                feat = out["warping_feat_output"]                 # tester.py:166-167, demo.py:241-243
        prev = data
        r = {"label": ops.argmax_channel(score)[0], "is_key": idx % interval == 0}
        if "score" in keep:
            r["score"] = score
        if "feat" in keep:
            r["feat"] = feat
        if "flow" in keep and flow is not None:
            r["flow"] = flow
        results.append(r)
    return results

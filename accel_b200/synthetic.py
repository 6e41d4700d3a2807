"""Seeded synthetic weights and frame sequences (there is no network for checkpoints or Cityscapes).

Weights are keyed by the reference's parameter names (netspec.param_spec); every tensor is drawn
from its own generator seeded with crc32(name) ^ seed, so the tensors two Accel versions share
(R101-DCN key net, FlowNet, L head) are identical across versions.  Distributions follow SURVEY.md
section 8(d), with the residual branches damped so activations stay O(1..30) through 33 bottleneck
units that have no live normalisation (inference BN uses the drawn moving statistics).

Frames mimic `lib/utils/image.py:transform` output: (1,3,H,W) fp32, RGB, minus the Cityscapes
pixel means of dff_deeplab_vid_demo.yaml:13-16.
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch

from .netspec import param_spec

PIXEL_MEANS_BGR = (103.06, 115.90, 123.15)       # experiments/dff_deeplab/cfgs/dff_deeplab_vid_demo.yaml:13-16


def bilinear_kernel(num_classes, factor=16):
    """MXNet `_init_bilinear` closed form (deeplab/symbols/resnet_v1_101_deeplab.py:820-828)."""
    k = 2 * factor
    f = math.ceil(k / 2.0)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    w1 = np.array([1 - abs(i / f - c) for i in range(k)], dtype=np.float32)
    return torch.from_numpy(np.broadcast_to(np.outer(w1, w1).astype(np.float32), (num_classes, 1, k, k)).copy())


def _draw(name, shape, kind, seed):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
    n = lambda std: torch.randn(shape, generator=g) * std
    u = lambda lo, hi: torch.rand(shape, generator=g) * (hi - lo) + lo
    residual_tail = name.endswith("_conv2_weight") and "_unit" in name
    if kind in ("conv", "flow_head", "offset", "score"):
        fan_in = shape[1] * shape[2] * shape[3]
        scale = {"flow_head": 0.5, "offset": 0.5}.get(kind, 0.3 if residual_tail else 1.0)
        if name in ("conv1_weight", "50_conv1_weight", "18_conv0_weight", "34_conv0_weight"):
            scale = 1.0 / 50.0                                       # pixel values are O(100), not O(1)
        elif name == "flow_conv1_weight":
            scale = 3.0                                              # FlowNet sees pixels / 255
        return n(scale * math.sqrt(2.0 / fan_in))
    if kind in ("deconv", "flow_up"):
        fan_in = shape[0] * (shape[2] // 2) * (shape[3] // 2)       # k4/s2: 2x2 taps reach each output
        return n(math.sqrt(2.0 / fan_in))
    if kind == "bias":
        return n(0.01)
    if kind == "gamma":
        damped = name.endswith("_branch2c_gamma") or ("bn5" in name and name.endswith("_branch2b_gamma")
                                                      and name[:3] in ("18_", "34_"))
        return u(0.1, 0.3) if damped else u(0.5, 1.5)
    if kind == "gamma_one":
        return torch.ones(shape)
    if kind == "beta":
        return n(0.1)
    if kind == "mean":
        return n(0.1)
    if kind == "var":
        return u(0.5, 1.5)
    if kind == "bilinear":
        return bilinear_kernel(shape[0], shape[2] // 2)
    if kind == "corr_score":                                       # 0.5*L + 0.5*R + noise
        w = n(0.05)
        c = shape[0]
        for i in range(c):
            w[i, i] += 0.5
            w[i, c + i] += 0.5
        return w
    if kind == "corr_feat":                                        # accel_101.py:276-289: [0 | I] + noise
        w = n(0.2 / math.sqrt(shape[1]))
        c = shape[0]
        idx = torch.arange(c)
        w[idx, c + idx, 0, 0] += 1.0
        return w
    raise ValueError(kind)


def make_params(version, seed=0):
    """{name: fp32 CPU tensor} for `version` in netspec.VERSIONS."""
    return {name: _draw(name, shape, kind, seed).float().contiguous()
            for name, (shape, kind) in param_spec(version).items()}


def _smooth_noise(h, w, gen, octaves=5):
    """Sum of bilinearly upsampled uniform noise at `octaves` scales -> (3,h,w) in [0,1].
    Interpolated and summed in float64: the uint8 frames derived from it must not depend on how many threads (or which
    vector ISA) the CPU interpolation kernel runs with -- in float32 a few pixels per frame flip at .5 boundaries
    between a 1-thread torchrun rank and a 16-thread single process, and with them the per-stream label CRCs."""
    img = torch.zeros(1, 3, h, w, dtype=torch.float64)
    amp_sum = 0.0
    for o in range(octaves):
        cells = 2 ** (o + 2)
        amp = 0.5 ** o
        coarse = torch.rand(1, 3, max(2, cells * h // w) + 1, cells + 1, generator=gen).double()
        img += amp * torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)
        amp_sum += amp
    return (img / amp_sum)[0]


def make_frames_u8(num_frames, height, width, stream=0, max_shift=8):
    """Synthetic BGR uint8 video (T,H,W,3): a smooth random texture translated along a smooth
    trajectory (|delta| <= max_shift px/frame) -- what cv2.imread would hand to the reference."""
    gen = torch.Generator().manual_seed(1000 + stream)
    margin = max_shift * num_frames + 2
    big = _smooth_noise(height + 2 * margin, width + 2 * margin, gen)
    big = (big * 255.0).clamp(0, 255)
    frames = []
    x = y = float(margin)
    ang = float(torch.rand(1, generator=gen)) * 2 * math.pi
    for t in range(num_frames):
        xi, yi = int(round(x)), int(round(y))
        f = big[:, yi:yi + height, xi:xi + width]
        frames.append(f.permute(1, 2, 0).round().to(torch.uint8))
        ang += (float(torch.rand(1, generator=gen)) - 0.5) * 0.8
        step = max_shift * (0.5 + 0.5 * float(torch.rand(1, generator=gen)))
        x = min(max(x + step * math.cos(ang), 0), 2 * margin)
        y = min(max(y + step * math.sin(ang), 0), 2 * margin)
    return torch.stack(frames)


def transform(frame_bgr_u8):
    """lib/utils/image.py:224-235 `transform`: (H,W,3) BGR -> (1,3,H,W) fp32 RGB minus pixel means.
    float64 arithmetic, one rounding to fp32 -- numpy's `im[:, :, 2 - i] - pixel_means[2 - i]` followed by
    mx.nd.array (demo.py:185); accel_preprocess is the device version of this."""
    im = frame_bgr_u8.to(torch.float64)
    out = torch.empty(1, 3, im.shape[0], im.shape[1], dtype=torch.float64)
    for i in range(3):
        out[0, i] = im[:, :, 2 - i] - PIXEL_MEANS_BGR[2 - i]
    return out.to(torch.float32)


def make_frames(num_frames, height, width, stream=0):
    """List of `transform`ed fp32 (1,3,H,W) tensors."""
    return [transform(f) for f in make_frames_u8(num_frames, height, width, stream)]

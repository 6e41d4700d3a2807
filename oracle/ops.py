"""MXNet operator semantics restated in PyTorch fp32 (CPU).  Test infrastructure -- see package doc.

Each function names the MXNet operator it stands for and the reference call sites that use it
(paths relative to /root/reference).  MXNet @ 62ecb60 itself is not vendored; rules tagged
[MXNet-ext] are the operator's published behaviour, stated explicitly here so a reader can check
them against upstream.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def convolution(x, weight, bias=None, stride=1, pad=0, dilate=1):
    """mx.symbol.Convolution -- cross-correlation, NCHW, weight (Cout, Cin, kh, kw) [MXNet-ext].

    Call sites: every conv of dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py (e.g. :52-83,
    :1754-1807) and the heads in accel_18.py:181-191."""
    return F.conv2d(x, weight, bias, stride=stride, padding=pad, dilation=dilate)


def deconvolution(x, weight, bias=None, stride=1, pad=0, num_group=1):
    """mx.symbol.Deconvolution -- transposed conv, weight (Cin, Cout/g, kh, kw), output size
    (i-1)*s + k - 2p [MXNet-ext].  Call sites: FlowNet refinement (...flownet_deeplab.py:1775-1799),
    `18_feat_upsampling` (accel_18.py:204-206), score `upsampling` (accel_18.py:193-195)."""
    return F.conv_transpose2d(x, weight, bias, stride=stride, padding=pad, groups=num_group)


def crop(x, like, offset):
    """mx.symbol.Crop(x, like, offset=(oy, ox)) = x[:, :, oy:oy+H_like, ox:ox+W_like] [MXNet-ext].
    Call sites: accel_18.py:197,227 (offset 8,8); ...flownet_deeplab.py:1776-1800 (offset 1,1)."""
    oy, ox = offset
    h, w = like.shape[2], like.shape[3]
    out = x[:, :, oy:oy + h, ox:ox + w]
    assert out.shape[2] == h and out.shape[3] == w, "Crop window leaves the source"
    return out


def batch_norm(x, gamma, beta, moving_mean, moving_var, eps, fix_gamma=False):
    """mx.symbol.BatchNorm at inference (use_global_stats / is_train=False):
    gamma * (x - mean) / sqrt(var + eps) + beta, gamma forced to 1 when fix_gamma [MXNet-ext].
    eps is 2e-5 in `residual_unit`/`resnet` (...flownet_deeplab.py:50-75,108,115) and self.eps=1e-5
    in the caffe-style nets (:23,136...)."""
    if fix_gamma:
        gamma = torch.ones_like(gamma)
    inv = gamma / torch.sqrt(moving_var + eps)
    return x * inv.view(1, -1, 1, 1) + (beta - moving_mean * inv).view(1, -1, 1, 1)


def pooling(x, kernel, stride, pad, pool_type, full):
    """mx.symbol.Pooling.  pooling_convention 'valid' -> floor, 'full' -> ceil output size with the
    window clipped at the border (max ignores the out-of-range taps) [MXNet-ext].  The 2x2/s2 'full'
    average pools (...flownet_deeplab.py:1753,1802) only ever see exact tiles on this path."""
    if pool_type == "max":
        return F.max_pool2d(x, kernel, stride, pad, ceil_mode=full)
    assert pool_type == "avg"
    assert pad == 0 and x.shape[2] % stride == 0 and x.shape[3] % stride == 0 and kernel == stride, \
        "average pooling is only defined here for exact tiles"
    return F.avg_pool2d(x, kernel, stride)


def leaky_relu(x, slope=0.1):
    """mx.symbol.LeakyReLU(act_type='leaky', slope=0.1): x > 0 ? x : slope * x."""
    return torch.where(x > 0, x, x * slope)


def grid_generator_warp(flow):
    """mx.sym.GridGenerator(transform_type='warp') (accel_18.py:174) [MXNet-ext]:
    grid[:,0] = (flow_x + x) / ((W-1)/2) - 1 ; grid[:,1] = (flow_y + y) / ((H-1)/2) - 1, fp32."""
    n, two, h, w = flow.shape
    assert two == 2
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, w).expand(n, h, w)
    ys = torch.arange(h, dtype=torch.float32).view(1, h, 1).expand(n, h, w)
    gx = (flow[:, 0] + xs) / np.float32((w - 1) / 2.0) - 1.0
    gy = (flow[:, 1] + ys) / np.float32((h - 1) / 2.0) - 1.0
    return torch.stack([gx, gy], dim=1)


def bilinear_sampler(data, grid):
    """mx.sym.BilinearSampler (accel_18.py:175) [MXNet-ext].  For each output pixel:
    x = (gx + 1) * (W-1) / 2, y = (gy + 1) * (H-1) / 2; the four integer neighbours contribute
    data * weight when the neighbour lies in [0, W-1] x [0, H-1] and 0 otherwise; weights are
    (1 - frac) for the top/left tap and frac for the other."""
    n, c, h, w = data.shape
    gx, gy = grid[:, 0], grid[:, 1]
    x = (gx + 1.0) * np.float32((w - 1) / 2.0)
    y = (gy + 1.0) * np.float32((h - 1) / 2.0)
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    wx0 = 1.0 - (x - x0)
    wy0 = 1.0 - (y - y0)
    out = torch.zeros(n, c, grid.shape[2], grid.shape[3], dtype=data.dtype)
    flat = data.reshape(n, c, h * w)
    for dy, wy in ((0, wy0), (1, 1.0 - wy0)):
        for dx, wx in ((0, wx0), (1, 1.0 - wx0)):
            xi = x0 + dx
            yi = y0 + dy
            ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long().view(n, 1, -1).expand(n, c, -1)
            v = torch.gather(flat, 2, idx).view(n, c, grid.shape[2], grid.shape[3])
            out = out + v * (wy * wx * ok).unsqueeze(1)
    return out


# DCNv1's "zero outside" rule is DISCONTINUOUS in the sampling position: a tap at p = -1e-6 contributes 0, at p = +1e-6 the
# full border value (same at p = H / W).  When DCN_TRACE is a list, every call appends a boolean (N, Ho, Wo) map of the
# output pixels that have a tap within DCN_TRACE_TAU of such a jump: there, any two implementations whose offset inputs
# differ by rounding noise may legitimately disagree by O(|x|) (tests/test_oracle_ops.py::test_dcn_border_rule_is_discontinuous).
DCN_TRACE = None
DCN_TRACE_TAU = 2e-4


def deformable_convolution(x, offset, weight, stride, pad, dilate, num_deformable_group):
    """mx.contrib.symbol.DeformableConvolution, DCNv1 (...flownet_deeplab.py:146-148,1235-1237)
    [MXNet-ext].  offset is (N, dg*2*kh*kw, Ho, Wo); inside deformable group g channel
    2*(i*kw+j) is dy and +1 is dx of tap (i, j).  Sample point p = o*stride - pad + tap*dilate + d.
    The sample is 0 unless 0 <= p_y < H and 0 <= p_x < W; inside, it is the bilinear blend of
    floor/floor+1 with the high index clamped to H-1 / W-1 (when floor >= H-1 both indices and the
    coordinate collapse to H-1).  Output = weight (Cout, Cin*kh*kw) @ sampled columns, no bias."""
    n, c, h, w = x.shape
    cout, cin, kh, kw = weight.shape
    assert cin == c
    ho = (h + 2 * pad - (dilate * (kh - 1) + 1)) // stride + 1
    wo = (w + 2 * pad - (dilate * (kw - 1) + 1)) // stride + 1
    assert offset.shape == (n, num_deformable_group * 2 * kh * kw, ho, wo), offset.shape
    cpg = c // num_deformable_group
    base_y = (torch.arange(ho, dtype=torch.float32) * stride - pad).view(1, ho, 1)
    base_x = (torch.arange(wo, dtype=torch.float32) * stride - pad).view(1, 1, wo)
    cols = torch.zeros(n, c, kh * kw, ho * wo, dtype=x.dtype)
    flat = x.reshape(n, c, h * w)
    critical = torch.zeros(n, ho, wo, dtype=torch.bool) if DCN_TRACE is not None else None
    for g in range(num_deformable_group):
        xs = flat[:, g * cpg:(g + 1) * cpg]
        for i in range(kh):
            for j in range(kw):
                k = i * kw + j
                oy = offset[:, g * 2 * kh * kw + 2 * k]
                ox = offset[:, g * 2 * kh * kw + 2 * k + 1]
                py = base_y + float(i * dilate) + oy
                px = base_x + float(j * dilate) + ox
                inside = (py >= 0) & (px >= 0) & (py < h) & (px < w)
                if critical is not None:
                    t = DCN_TRACE_TAU
                    near_y = (py.abs() < t) | ((py - h).abs() < t)
                    near_x = (px.abs() < t) | ((px - w).abs() < t)
                    critical |= (near_y & (px > -t) & (px < w + t)) | (near_x & (py > -t) & (py < h + t))
                y0 = torch.floor(py)
                x0 = torch.floor(px)
                top = y0 >= h - 1
                left = x0 >= w - 1
                y0 = torch.where(top, torch.full_like(y0, h - 1), y0)
                x0 = torch.where(left, torch.full_like(x0, w - 1), x0)
                y1 = torch.where(top, y0, y0 + 1)
                x1 = torch.where(left, x0, x0 + 1)
                pyc = torch.where(top, y0, py)
                pxc = torch.where(left, x0, px)
                ly = pyc - y0
                lx = pxc - x0
                hy = 1.0 - ly
                hx = 1.0 - lx
                acc = torch.zeros(n, cpg, ho * wo, dtype=x.dtype)
                for yy, xx, wgt in ((y0, x0, hy * hx), (y0, x1, hy * lx), (y1, x0, ly * hx), (y1, x1, ly * lx)):
                    idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).long().view(n, 1, -1).expand(n, cpg, -1)
                    acc = acc + torch.gather(xs, 2, idx) * (wgt * inside).view(n, 1, -1)
                cols[:, g * cpg:(g + 1) * cpg, k] = acc
    if critical is not None:
        DCN_TRACE.append(critical)
    out = torch.einsum("ok,nkp->nop", weight.reshape(cout, c * kh * kw), cols.reshape(n, c * kh * kw, ho * wo))
    return out.view(n, cout, ho, wo)


def bilinear_upsampling_weight(num_classes, factor=16):
    """Fixed `upsampling_weight` (num_classes, 1, 2f, 2f) as the reference initialises it
    (deeplab/symbols/resnet_v1_101_deeplab.py:820-828): f = ceil(k/2), c = (2f-1-f%2)/(2f),
    w[y,x] = (1-|x/f-c|)(1-|y/f-c|); lr_mult 0 keeps it fixed."""
    k = 2 * factor
    f = math.ceil(k / 2.0)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    w1 = np.array([1 - abs(i / f - c) for i in range(k)], dtype=np.float32)
    w2 = np.outer(w1, w1).astype(np.float32)
    return torch.from_numpy(np.broadcast_to(w2, (num_classes, 1, k, k)).copy())


def argmax_channel(score):
    """mx.ndarray.argmax(score, axis=1) -> np.uint8 (dff_deeplab/demo.py:238,245,252); ties go to
    the lowest class index [MXNet-ext].  Integer result: the bit-exact parity target."""
    s = score.detach().cpu().numpy()
    return np.argmax(s, axis=1).astype(np.uint8)  # numpy returns the first maximal index


def fast_hist(pred, label, n):
    """Confusion matrix, dff_deeplab/demo.py:50-53: rows = label, cols = prediction, labels >= n
    (ignore 255) dropped."""
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    label = np.asarray(label).reshape(-1).astype(np.int64)
    k = (label >= 0) & (label < n)
    return np.bincount(n * label[k] + pred[k], minlength=n * n).reshape(n, n)


def per_class_iu(hist):
    """dff_deeplab/demo.py:55-56."""
    hist = hist.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))

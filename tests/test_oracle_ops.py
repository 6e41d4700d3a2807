"""The oracle's MXNet-operator restatements against independent formulations (CPU only).

The reference pins nothing at op granularity, so each [MXNet-ext] rule in
oracle/ops.py is cross-checked against a second, independently written statement of the same rule.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops


def test_bilinear_sampler_matches_grid_sample():
    g = torch.Generator().manual_seed(1)
    data = torch.randn(2, 5, 9, 13, generator=g)
    flow = torch.randn(2, 2, 9, 13, generator=g) * 3.0          # some samples leave the map
    grid = ops.grid_generator_warp(flow)
    mine = ops.bilinear_sampler(data, grid)
    ref = F.grid_sample(data, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros", align_corners=True)
    assert torch.allclose(mine, ref, atol=2e-6)


def test_warp_zero_flow_is_identity_and_integer_flow_is_shift():
    data = torch.arange(2 * 6 * 8, dtype=torch.float32).view(1, 2, 6, 8)
    zero = torch.zeros(1, 2, 6, 8)
    assert torch.allclose(ops.bilinear_sampler(data, ops.grid_generator_warp(zero)), data, atol=1e-4)
    flow = zero.clone()
    flow[:, 0] = 2.0                                            # sample two pixels to the right
    flow[:, 1] = -1.0                                           # and one row up
    out = ops.bilinear_sampler(data, ops.grid_generator_warp(flow))
    expect = torch.zeros_like(data)
    expect[:, :, 1:, :6] = data[:, :, :5, 2:]
    assert torch.allclose(out, expect, atol=1e-3)


def test_deformable_conv_zero_offset_is_dilated_conv():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 8, 10, 12, generator=g)
    w = torch.randn(6, 8, 3, 3, generator=g)
    off = torch.zeros(1, 4 * 18, 10, 12)
    out = ops.deformable_convolution(x, off, w, 1, 2, 2, 4)
    assert torch.allclose(out, F.conv2d(x, w, None, 1, 2, 2), atol=1e-4)


def test_deformable_conv_matches_torchvision_on_interior_samples():
    tv = pytest.importorskip("torchvision.ops")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 8, 16, 16, generator=g)
    w = torch.randn(4, 8, 3, 3, generator=g)
    # offsets below 1 px; evaluate only outputs whose taps stay >= 1 px inside the map, where the
    # MXNet rule and torchvision's zero-blend border rule coincide.
    off = (torch.rand(1, 2 * 18, 16, 16, generator=g) - 0.5) * 1.6
    mine = ops.deformable_convolution(x, off, w, 1, 2, 2, 2)
    ref = tv.deform_conv2d(x, off, w, None, stride=1, padding=2, dilation=2)
    assert torch.allclose(mine[:, :, 4:-4, 4:-4], ref[:, :, 4:-4, 4:-4], atol=1e-4)


def test_deformable_conv_border_rule():
    # a sample at y in (-1, 0) is zero (torchvision would blend with zero); a sample at
    # y in (H-1, H) collapses onto row H-1 with weight 1.
    x = torch.ones(1, 1, 4, 4)
    w = torch.zeros(1, 1, 3, 3)
    w[0, 0, 1, 1] = 1.0                                         # only the centre tap
    off = torch.zeros(1, 18, 4, 4)
    off[0, 2 * 4, 0, :] = -0.5                                  # centre tap dy at output row 0
    off[0, 2 * 4, 3, :] = +0.5                                  # centre tap dy at output row 3
    out = ops.deformable_convolution(x, off, w, 1, 1, 1, 1)
    assert torch.all(out[0, 0, 0] == 0)
    assert torch.all(out[0, 0, 3] == 1)
    assert torch.all(out[0, 0, 1:3] == 1)


def test_pooling_conventions():
    x = torch.randn(1, 3, 16, 32)
    assert ops.pooling(x, 3, 2, 0, "max", full=True).shape == (1, 3, 8, 16)       # ceil((16-3)/2)+1
    assert ops.pooling(x, 3, 2, 1, "max", full=False).shape == (1, 3, 8, 16)      # floor((16+2-3)/2)+1
    y = ops.pooling(x, 3, 2, 0, "max", full=True)
    assert y[0, 0, 7, 15] == x[0, 0, 14:16, 30:32].max()                          # clipped window
    assert torch.allclose(ops.pooling(x, 2, 2, 0, "avg", full=True),
                          x.view(1, 3, 8, 2, 16, 2).mean(dim=(3, 5)), atol=1e-6)


def test_flownet_deconv_crop_equals_padded_transposed_conv():
    # Deconvolution(k4,s2,p0) + Crop(offset 1,1) == conv_transpose2d(padding=1)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 5, 6, 7, generator=g)
    w = torch.randn(5, 3, 4, 4, generator=g)
    b = torch.randn(3, generator=g)
    skip = torch.zeros(1, 1, 12, 14)
    a = ops.crop(ops.deconvolution(x, w, b, 2, 0), skip, (1, 1))
    assert torch.allclose(a, F.conv_transpose2d(x, w, b, stride=2, padding=1), atol=1e-5)


def test_score_upsampling_is_four_tap_bilinear():
    # grouped 32x32/s16 deconv + Crop(8,8): out[Y] = sum_i in[i] * w1[Y + 8 - 16 i], two taps per axis
    g = torch.Generator().manual_seed(5)
    s = torch.randn(1, 19, 4, 6, generator=g)
    up = ops.deconvolution(s, ops.bilinear_upsampling_weight(19), None, 16, 0, num_group=19)
    up = up[:, :, 8:8 + 64, 8:8 + 96]
    Y, X = 37, 50
    acc = torch.zeros(19)
    for i in range(4):
        for j in range(6):
            ky, kx = Y + 8 - 16 * i, X + 8 - 16 * j
            if 0 <= ky < 32 and 0 <= kx < 32:
                acc += s[0, :, i, j] * ops.bilinear_upsampling_weight(1)[0, 0, ky, kx]
    assert torch.allclose(up[0, :, Y, X], acc, atol=1e-5)
    k1 = np.array([1 - abs(i / 16.0 - 31 / 32.0) for i in range(32)], dtype=np.float32)
    assert np.allclose(ops.bilinear_upsampling_weight(2)[1, 0].numpy(), np.outer(k1, k1))


def test_argmax_ties_take_lowest_index_and_cast():
    s = torch.zeros(1, 19, 2, 2)
    s[0, 5, 0, 0] = 1.0
    s[0, 7, 0, 0] = 1.0
    lab = ops.argmax_channel(s)
    assert lab.dtype == np.uint8 and lab.shape == (1, 2, 2)
    assert lab[0, 0, 0] == 5 and lab[0, 1, 1] == 0


def test_batchnorm_fix_gamma():
    x = torch.randn(1, 3, 4, 4)
    gamma = torch.tensor([2.0, 3.0, 4.0])
    beta, mean, var = torch.randn(3), torch.randn(3), torch.rand(3) + 0.5
    a = ops.batch_norm(x, gamma, beta, mean, var, 2e-5, fix_gamma=True)
    b = (x - mean.view(1, 3, 1, 1)) / torch.sqrt(var.view(1, 3, 1, 1) + 2e-5) + beta.view(1, 3, 1, 1)
    assert torch.allclose(a, b, atol=1e-5)


def test_fast_hist_and_iu():
    pred = np.array([0, 1, 1, 2, 2, 2])
    label = np.array([0, 1, 2, 2, 255, 2])
    h = ops.fast_hist(pred, label, 3)
    assert h.tolist() == [[1, 0, 0], [0, 1, 0], [0, 1, 2]]
    iu = ops.per_class_iu(h)
    assert np.allclose(iu, [1.0, 0.5, 2.0 / 3.0])


def test_dcn_border_rule_is_discontinuous():
    """DCNv1's `0 unless 0 <= p < H` rule jumps at the image border: shifting one tap's offset by 2e-6 across p = 0 changes
    the output by O(|x|), and ops.DCN_TRACE flags exactly that output pixel.  This is why graph-level parity is stated
    away from `border-critical` deformable samples (tests/parity_util.py)."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 8, 6, 7, generator=g) + 1.0
    w = torch.rand(4, 8, 3, 3, generator=g)
    off = torch.full((1, 18, 6, 7), 0.37)                      # generic positions: no other tap sits on a border
    # tap (i=1, j=0) of output pixel (2, 1): base x = 1 - 2 + 0 = -1; dx = 1 +- 1e-6 puts it just inside / outside x = 0
    outs = []
    for d in (1.0 + 1e-6, 1.0 - 1e-6):
        o = off.clone()
        o[0, 2 * 3 + 1, 2, 1] = d
        ops.DCN_TRACE = []
        outs.append(ops.deformable_convolution(x, o, w, 1, 2, 2, 1))
        trace = ops.DCN_TRACE
        ops.DCN_TRACE = None
        assert trace[0].sum().item() == 1 and bool(trace[0][0, 2, 1])
    diff = (outs[0] - outs[1]).abs()
    assert diff[0, :, 2, 1].max().item() > 0.5                 # an O(1) jump for a 2e-6 change of the offset ...
    diff[0, :, 2, 1] = 0
    assert diff.max().item() == 0.0                             # ... at that pixel only

"""Operator-level parity of the CUDA kernels against the CPU oracle, through the C ABI (-m gpu)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from accel_b200 import engine as E
from oracle import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ----------------------------------------------------------------------------- warp (a2)
@pytest.mark.parametrize("c,h,w", [(64, 8, 16), (2048, 8, 16), (96, 13, 37), (2048, 64, 128)])
def test_warp_is_bit_exact_vs_oracle(c, h, w):
    feat = _rand(1, c, h, w, seed=1)
    flow = _rand(1, 2, h, w, seed=2, scale=3.0)                 # several samples leave the map
    flow[0, :, 0, 0] = torch.tensor([1e12, -1e12])              # absurd flow: every tap out of range -> 0
    ref = ops.bilinear_sampler(feat, ops.grid_generator_warp(flow))
    out = E.warp(feat.to(DEV), flow.to(DEV)).cpu()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("c,h,w,scale", [(2048, 64, 128, 0.7), (20, 64, 128, 1.5), (44, 32, 64, 0.5), (24, 24, 120, 1.0),
                                         (16, 128, 256, 2.0), (8, 4, 1024, 0.5)])
def test_warp_staged_rows_bit_exact(c, h, w, scale):
    """Smooth FlowNet-like fields (the shared-memory staged path: narrow source-row band per CTA), channel
    counts that are not a multiple of the 8-channel stage, widths that are not a power of two."""
    feat = _rand(1, c, h, w, seed=4)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    flow = torch.stack([scale * (3.0 * torch.sin(yy / 9.0) + 0.01 * xx - 2.0), scale * 2.0 * torch.cos(xx / 13.0 + yy / 7.0)])[None]
    flow = (flow + _rand(1, 2, h, w, seed=5, scale=0.05)).contiguous()
    ref = ops.bilinear_sampler(feat, ops.grid_generator_warp(flow))
    out = E.warp(feat.to(DEV), flow.to(DEV)).cpu()
    assert torch.equal(out, ref)


def test_warp_mixed_band_heights():
    """Half of the rows see a smooth field (staged CTAs), the other half a wild one (gather CTAs) in one launch."""
    c, h, w = 40, 64, 128
    feat = _rand(1, c, h, w, seed=6)
    flow = _rand(1, 2, h, w, seed=7, scale=0.8)
    flow[:, :, 32:] = _rand(1, 2, 32, w, seed=8, scale=25.0)
    ref = ops.bilinear_sampler(feat, ops.grid_generator_warp(flow))
    out = E.warp(feat.to(DEV), flow.to(DEV)).cpu()
    assert torch.equal(out, ref)


def _split_ref(ref):
    """hi = rn16(v), lo = rn16(v - hi) of an fp32 NCHW tensor, as (H,W,C) planes (split_pair, common.cuh)."""
    hi = ref[0].permute(1, 2, 0).contiguous().to(torch.float16)
    lo = (ref[0].permute(1, 2, 0) - hi.float()).to(torch.float16)
    return hi, lo


@pytest.mark.parametrize("c,h,w,kind", [(2048, 64, 128, "smooth"), (64, 64, 128, "wild"), (96, 32, 64, "mixed"), (32, 8, 16, "smooth"),
                                        (64, 13, 36, "smooth"), (2048, 64, 128, "noisy")])
@pytest.mark.parametrize("per_sm", ["2", "3"])
def test_warp_split_both_outputs_bit_exact(c, h, w, kind, per_sm, monkeypatch):
    """accel_warp_split: the fp32 NCHW warp equals the oracle bit for bit and the split-fp16 NHWC head operand equals the
    split of that result, for fields the ring holds (smooth), fields it does not (wild: gather CTAs), both in one launch,
    and a shape the fused kernel declines (13x36: warp + layout pass)."""
    monkeypatch.setenv("ACCEL_WARP_FUSED_PER_SM", per_sm)
    feat = _rand(1, c, h, w, seed=11)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    smooth = torch.stack([3.0 * torch.sin(yy / 9.0) + 0.01 * xx - 2.0, 2.0 * torch.cos(xx / 13.0 + yy / 7.0)])[None]
    if kind == "smooth":
        flow = smooth + _rand(1, 2, h, w, seed=12, scale=0.05)
    elif kind == "noisy":
        flow = smooth + _rand(1, 2, h, w, seed=12, scale=0.6)
    elif kind == "wild":
        flow = _rand(1, 2, h, w, seed=12, scale=25.0)
    else:
        flow = smooth.clone()
        flow[:, :, h // 2:] = _rand(1, 2, h - h // 2, w, seed=13, scale=25.0)
    flow = flow.contiguous()
    ref = ops.bilinear_sampler(feat, ops.grid_generator_warp(flow))
    out, hi, lo = E.warp_split(feat.to(DEV), flow.to(DEV))
    assert torch.equal(out.cpu(), ref)
    rhi, rlo = _split_ref(ref)
    assert torch.equal(hi.cpu(), rhi)
    assert torch.equal(lo.cpu(), rlo)


@pytest.mark.parametrize("gather", ["0", "1"])
def test_warp_out_of_range_taps_are_skipped_by_value(gather, monkeypatch):
    """MXNet's BilinearSampler skips out-of-range taps; a zero WEIGHT on a clamped neighbour is not the same thing when
    that neighbour holds Inf (0 * Inf = NaN).  A pixel whose sample lies wholly outside the map must read 0 even if the
    feature at its own location is Inf."""
    monkeypatch.setenv("ACCEL_WARP_GATHER", gather)
    c, h, w = 64, 16, 128
    feat = _rand(1, c, h, w, seed=40)
    feat[0, :, 5, 7] = float("inf")
    flow = torch.zeros(1, 2, h, w)
    flow[0, 0, 5, 7] = 1000.0                       # (5, 7) samples far outside; every other pixel is the identity
    flow[0, 0, 5, 6] = 0.5                          # its left neighbour blends (5,6) and (5,7): Inf, legitimately
    out = E.warp(feat.to(DEV), flow.to(DEV)).cpu()
    ref = ops.bilinear_sampler(feat, ops.grid_generator_warp(flow))
    assert torch.isfinite(ref[0, :, 5, 7]).all() and (ref[0, :, 5, 7] == 0).all()
    assert (out[0, :, 5, 7] == 0).all()
    assert torch.isinf(out[0, :, 5, 6]).all()
    # everywhere else: the oracle's result bit for bit, including the NaNs an in-range tap of weight +0 next to the Inf
    # produces there too (0 * Inf)
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isnan(out), torch.isnan(ref)) and torch.equal(torch.isinf(out), torch.isinf(ref))
    assert torch.equal(out[fin], ref[fin])


def test_warp_zero_flow_identity_and_integer_shift():
    feat = _rand(1, 32, 16, 32, seed=3)
    zero = torch.zeros(1, 2, 16, 32)
    assert torch.allclose(E.warp(feat.to(DEV), zero.to(DEV)).cpu(), feat, atol=1e-5)
    flow = zero.clone()
    flow[:, 0] = 3.0
    out = E.warp(feat.to(DEV), flow.to(DEV)).cpu()
    assert torch.allclose(out[..., :29], feat[..., 3:], atol=1e-4) and torch.all(out[..., 29:] == 0)


# ----------------------------------------------------------------------------- tail (a3/a10/a12)
def _oracle_tail(a, b, wc, bc):
    k = a.shape[1]
    up = lambda s: ops.deconvolution(s, ops.bilinear_upsampling_weight(k), None, 16, 0, num_group=k)[
        :, :, 8:8 + 16 * s.shape[2], 8:8 + 16 * s.shape[3]]
    if wc is None:
        return up(a)
    return ops.convolution(torch.cat([up(a), up(b)], dim=1), wc, bc)


@pytest.mark.parametrize("h,w", [(8, 16), (5, 7), (64, 128)])
@pytest.mark.parametrize("fused", [False, True])
def test_tail_scores_and_labels(h, w, fused):
    k = 19
    a, b = _rand(1, k, h, w, seed=4, scale=3.0), _rand(1, k, h, w, seed=5, scale=3.0)
    wc = (_rand(k, 2 * k, 1, 1, seed=6, scale=0.05) + torch.cat([torch.eye(k), torch.eye(k)], 1).view(k, 2 * k, 1, 1) * 0.5) if fused else None
    bc = _rand(k, seed=7, scale=0.1) if fused else None
    ref = _oracle_tail(a, b, wc, bc)
    label, full = E.fuse_argmax(a.to(DEV), b.to(DEV) if fused else None, wc.to(DEV) if fused else None,
                                bc.to(DEV) if fused else None, want_scores=True)
    full, label = full.cpu(), label.cpu().numpy()
    assert (full - ref).abs().max().item() < 2e-5
    # integer work: the label map is bit-exactly the lowest-index argmax of the emitted score volume
    assert np.array_equal(label, ops.argmax_channel(full)[0])
    # and agrees with the oracle's labels wherever the oracle's top-2 margin exceeds the fp32 noise
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[0].numpy()
    assert np.array_equal(label[margin > 1e-4], ops.argmax_channel(ref)[0][margin > 1e-4])


def test_tail_ties_take_lowest_class():
    s = torch.zeros(1, 19, 4, 4)
    s[0, 7] = 1.0
    s[0, 11] = 1.0
    label = E.fuse_argmax(s.to(DEV)).cpu().numpy()
    assert label.shape == (64, 64) and label.dtype == np.uint8
    assert np.all(label == 7)
    label0 = E.fuse_argmax(torch.zeros(1, 19, 4, 4).to(DEV)).cpu().numpy()
    assert np.all(label0 == 0)


# ----------------------------------------------------------------------------- task head (a3 / a7)
@pytest.mark.parametrize("cin,mid,k,h,w", [(2048, 1024, 19, 8, 16), (512, 256, 19, 5, 7)])
def test_head_matches_torch(cin, mid, k, h, w):
    feat = F.relu(_rand(1, cin, h, w, seed=50))
    w1, b1 = _rand(mid, cin, 1, 1, seed=51, scale=(2.0 / cin) ** 0.5), _rand(mid, seed=52, scale=0.1)
    w2, b2 = _rand(k, mid, 1, 1, seed=53, scale=(2.0 / mid) ** 0.5), _rand(k, seed=54, scale=0.1)
    ref = F.conv2d(F.relu(F.conv2d(feat, w1, b1)), w2, b2)
    out = E.head(feat.to(DEV), w1, b1, w2, b2).cpu()
    assert out.shape == ref.shape and (out - ref).abs().max().item() < _tol(ref)


# ----------------------------------------------------------------------------- conv engines (a1, a3-a9)
CONV_CASES = [
    # cin, cout, h, w, k, stride, pad, dil
    (64, 64, 16, 32, 3, 1, 1, 1),
    (64, 128, 16, 32, 3, 2, 1, 1),
    (64, 128, 16, 32, 1, 2, 0, 1),
    (128, 256, 16, 16, 5, 2, 2, 1),
    (256, 72, 8, 16, 3, 1, 2, 2),
    (194, 2, 8, 16, 3, 1, 1, 1),
    (1026, 2, 2, 4, 3, 1, 1, 1),
    (2048, 1024, 8, 16, 1, 1, 0, 1),
    (1024, 19, 8, 16, 1, 1, 0, 1),
    (512, 512, 2, 4, 3, 1, 1, 1),
    (386, 64, 9, 11, 3, 1, 1, 1),
]


def _tol(ref):
    return 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_layer_matches_torch(case, engine):
    cin, cout, h, w, k, s, p, d = case
    x = _rand(1, cin, h, w, seed=10)
    wt = _rand(cout, cin, k, k, seed=11, scale=(2.0 / (cin * k * k)) ** 0.5)
    scale, shift = torch.rand(cout, generator=torch.Generator().manual_seed(12)) + 0.5, _rand(cout, seed=13, scale=0.1)
    ref = F.conv2d(x, wt, None, s, p, d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = _rand(*ref.shape, seed=14)
    ref = F.relu(ref + res)
    out = E.conv_layer(x.to(DEV), wt, "conv", s, p, d, scale, shift, act=1, residual=res.to(DEV), engine=engine).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


# CTA-pair kernel (conv_tc2_kernel, cta_group::2: 256-pixel x 128-channel tiles, each SM stages half of the weight rows):
# forced on for every shape the planner admits -- odd tile counts (a phantom tile in the last pair), partial tiles,
# stride 2, few / many output channels, long K (two accumulation chains as K halves), a residual.
PAIR_CASES = CONV_CASES + [(256, 256, 64, 128, 3, 1, 1, 1), (1024, 256, 24, 128, 1, 1, 0, 1), (512, 2048, 5, 128, 1, 1, 0, 1),
                           (2048, 512, 64, 128, 1, 1, 0, 1)]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_layer_cta_pairs(case, monkeypatch):
    monkeypatch.setenv("ACCEL_TC_PAIR", "1")
    cin, cout, h, w, k, s, p, d = case
    x = _rand(1, cin, h, w, seed=50)
    wt = _rand(cout, cin, k, k, seed=51, scale=(2.0 / (cin * k * k)) ** 0.5)
    scale, shift = torch.rand(cout, generator=torch.Generator().manual_seed(52)) + 0.5, _rand(cout, seed=53, scale=0.1)
    ref = F.conv2d(x, wt, None, s, p, d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = _rand(*ref.shape, seed=54)
    ref = F.relu(ref + res)
    out = E.conv_layer(x.to(DEV), wt, "conv", s, p, d, scale, shift, act=1, residual=res.to(DEV), engine=2).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_layer_weight_multicast(case, monkeypatch):
    """ACCEL_TC_MCAST=1: clusters of two CTAs on two M tiles of one N tile, each CTA loads one plane of the weight tile and
    multicasts it into both rings (TcParams::mcast) -- forced on for every shape the planner admits."""
    monkeypatch.setenv("ACCEL_TC_MCAST", "1")
    cin, cout, h, w, k, s, p, d = case
    x = _rand(1, cin, h, w, seed=60)
    wt = _rand(cout, cin, k, k, seed=61, scale=(2.0 / (cin * k * k)) ** 0.5)
    scale, shift = torch.rand(cout, generator=torch.Generator().manual_seed(62)) + 0.5, _rand(cout, seed=63, scale=0.1)
    ref = F.conv2d(x, wt, None, s, p, d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = _rand(*ref.shape, seed=64)
    ref = F.relu(ref + res)
    out = E.conv_layer(x.to(DEV), wt, "conv", s, p, d, scale, shift, act=1, residual=res.to(DEV), engine=2).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("knob", ["ACCEL_TC_EPI1", "ACCEL_TC_EPIW16"])
@pytest.mark.parametrize("case", [(64, 256, 16, 128, 1, 1, 0, 1), (256, 1024, 9, 128, 1, 1, 0, 1), (512, 2048, 4, 128, 1, 1, 0, 1),
                                  (128, 512, 32, 64, 1, 1, 0, 1), (256, 128, 8, 128, 1, 1, 0, 1)])
def test_conv_layer_half_chunk_tma_epilogues(case, knob, monkeypatch):
    """The two half-chunk variants of the TMA epilogue forced on for short-K layers with a residual: 16 epilogue warps
    (conv_tc_kernel<..., 16>) and one chunk set with a third operand stage (conv_tc_kernel<..., 8, 1>, exactly 227 KB)."""
    monkeypatch.setenv(knob, "1")
    cin, cout, h, w, k, s, p, d = case
    x = _rand(1, cin, h, w, seed=70)
    wt = _rand(cout, cin, k, k, seed=71, scale=(2.0 / (cin * k * k)) ** 0.5)
    scale, shift = torch.rand(cout, generator=torch.Generator().manual_seed(72)) + 0.5, _rand(cout, seed=73, scale=0.1)
    ref = F.conv2d(x, wt, None, s, p, d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = _rand(*ref.shape, seed=74)
    ref = F.relu(ref + res)
    out = E.conv_layer(x.to(DEV), wt, "conv", s, p, d, scale, shift, act=1, residual=res.to(DEV), engine=2).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


# Wide maps (W >= 128): one-row tiles, where the taps of a filter row share one TMA slab (TcParams::aslab, tap-shifted
# shared-memory descriptors) -- incl. a right border that is not a multiple of the tile, dilation 2 (the offset convs of
# res5), few and many output channels, a long K (two accumulation chains) and a transposed-conv phase set.
WIDE_CASES = [
    # cin, cout, h, w, k, pad, dil
    (64, 64, 6, 128, 3, 1, 1),
    (256, 256, 4, 128, 3, 1, 1),
    (128, 128, 3, 256, 3, 1, 1),
    (96, 40, 5, 160, 3, 1, 1),
    (512, 72, 3, 128, 3, 2, 2),
    (194, 2, 4, 256, 3, 1, 1),
    (512, 512, 2, 128, 3, 1, 1),
]


@pytest.mark.parametrize("slab", ["0", "1"])
@pytest.mark.parametrize("case", WIDE_CASES)
def test_conv_layer_wide_rows_slab_reuse(case, slab, monkeypatch):
    monkeypatch.setenv("ACCEL_TC_ASLAB", slab)          # 1 = the tap-shifted slab mode (measured slower: off by default)
    cin, cout, h, w, k, p, d = case
    x = _rand(1, cin, h, w, seed=20)
    wt = _rand(cout, cin, k, k, seed=21, scale=(2.0 / (cin * k * k)) ** 0.5)
    scale, shift = torch.rand(cout, generator=torch.Generator().manual_seed(22)) + 0.5, _rand(cout, seed=23, scale=0.1)
    ref = F.conv2d(x, wt, None, 1, p, d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = _rand(*ref.shape, seed=24)
    ref = F.relu(ref + res)
    out = E.conv_layer(x.to(DEV), wt, "conv", 1, p, d, scale, shift, act=1, residual=res.to(DEV), engine=2).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("slab", ["0", "1"])
@pytest.mark.parametrize("cin,cout,h,w", [(256, 128, 3, 128), (130, 64, 2, 192)])
def test_deconv_layer_wide_rows_slab_reuse(cin, cout, h, w, slab, monkeypatch):
    monkeypatch.setenv("ACCEL_TC_ASLAB", slab)
    x = _rand(1, cin, h, w, seed=30)
    wt = _rand(cin, cout, 4, 4, seed=31, scale=(2.0 / (cin * 4)) ** 0.5)
    ref = F.conv_transpose2d(x, wt, None, 2, 1)
    out = E.conv_layer(x.to(DEV), wt, "deconv", engine=2).cpu()
    assert out.shape == ref.shape and (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("case", [(2048, 1024, 8, 16, 1, 1, 0, 1), (512, 512, 2, 4, 3, 1, 1, 1), (1026, 2, 2, 4, 3, 1, 1, 1)])
def test_fused_splitk_tail_is_bit_identical(case, monkeypatch):
    """ACCEL_TC_FUSED_SPLITK=1 (the CTA that delivers a tile's last partial slab sums the slabs and runs the epilogue
    itself; measured slower and off by default) must give exactly the bits of the two-launch split-K path:
    same slabs, same summation order, same epilogue."""
    cin, cout, h, w, k, s, p, d = case
    x = _rand(1, cin, h, w, seed=40)
    wt = _rand(cout, cin, k, k, seed=41, scale=(2.0 / (cin * k * k)) ** 0.5)
    scale, shift = torch.rand(cout, generator=torch.Generator().manual_seed(42)) + 0.5, _rand(cout, seed=43, scale=0.1)
    res = _rand(1, cout, (h + 2 * p - d * (k - 1) - 1) // s + 1, (w + 2 * p - d * (k - 1) - 1) // s + 1, seed=44)
    outs = []
    for fused in ("0", "1"):
        monkeypatch.setenv("ACCEL_TC_FUSED_SPLITK", fused)
        monkeypatch.setenv("ACCEL_TC_SPLITS", "4")
        outs.append(E.conv_layer(x.to(DEV), wt, "conv", s, p, d, scale, shift, act=1, residual=res.to(DEV), engine=2).cpu())
    assert torch.equal(outs[0], outs[1])
    ref = F.relu(F.conv2d(x, wt, None, s, p, d) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    assert (outs[1] - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("cin,cout,h,w", [(1024, 512, 2, 4), (386, 64, 8, 16), (512, 2048, 4, 8)])
def test_deconv_layer_matches_torch(cin, cout, h, w, engine):
    x = _rand(1, cin, h, w, seed=20)
    wt = _rand(cin, cout, 4, 4, seed=21, scale=(2.0 / (cin * 4)) ** 0.5)
    shift = _rand(cout, seed=22, scale=0.1)
    ref = ops.leaky_relu(F.conv_transpose2d(x, wt, shift, stride=2, padding=1))
    out = E.conv_layer(x.to(DEV), wt, "deconv", shift=shift, act=2, engine=engine).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("dg", [1, 4])
def test_deformable_layer_matches_oracle(dg, engine):
    cin, cout, h, w = 64, 64, 12, 20
    x = _rand(1, cin, h, w, seed=30)
    wt = _rand(cout, cin, 3, 3, seed=31, scale=(2.0 / (cin * 9)) ** 0.5)
    off = _rand(1, dg * 18, h, w, seed=32, scale=1.5)           # some taps leave the map / hit the border ring
    ref = ops.deformable_convolution(x, off, wt, 1, 2, 2, dg)
    out = E.conv_layer(x.to(DEV), wt, "deform", 1, 2, 2, offset=off.to(DEV), deform_groups=dg, engine=engine).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("dg,cin,h,w,scale", [(1, 512, 64, 128, 1.0), (4, 512, 32, 64, 1.5), (4, 128, 19, 45, 6.0), (1, 64, 8, 16, 0.0)])
def test_deformable_layer_staged_shapes(dg, cin, h, w, scale):
    """Shapes of the real res5 layers (TMA-staged im2col: tile + halo in shared memory), a map that is not a
    multiple of the 8x16 tile with offsets large enough to leave the staged halo, and zero offsets."""
    cout = 64
    x = _rand(1, cin, h, w, seed=33)
    wt = _rand(cout, cin, 3, 3, seed=34, scale=(2.0 / (cin * 9)) ** 0.5)
    off = _rand(1, dg * 18, h, w, seed=35, scale=scale)
    ref = ops.deformable_convolution(x, off, wt, 1, 2, 2, dg)
    out = E.conv_layer(x.to(DEV), wt, "deform", 1, 2, 2, offset=off.to(DEV), deform_groups=dg, engine=2).cpu()
    assert (out - ref).abs().max().item() < _tol(ref)


# ----------------------------------------------------------------------------- 7x7/s2 stems (a1, a4, a8, a9)
@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("h,w", [(64, 128), (128, 256), (72, 200)])
def test_stem_matches_torch(h, w, engine):
    x = _rand(1, 3, h, w, seed=40, scale=60.0)                 # mean-subtracted pixel range
    wt = _rand(64, 3, 7, 7, seed=41, scale=0.02)
    scale, shift = torch.rand(64, generator=torch.Generator().manual_seed(42)) + 0.5, _rand(64, seed=43, scale=0.1)
    ref = F.relu(F.conv2d(x, wt, None, 2, 3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    out = E.conv_layer(x.to(DEV), wt, "stem", 2, 3, 1, scale, shift, act=1, engine=engine).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() < _tol(ref)


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("h,w", [(128, 256), (256, 384)])
def test_flownet_stem_matches_torch(h, w, engine):
    a, b = _rand(1, 3, h, w, seed=44, scale=60.0), _rand(1, 3, h, w, seed=45, scale=60.0)
    wt = _rand(64, 6, 7, 7, seed=46, scale=0.3)
    shift = _rand(64, seed=47, scale=0.1)
    x = F.avg_pool2d(torch.cat([a, b], 1) / 255.0, 2, 2)
    ref = ops.leaky_relu(F.conv2d(x, wt, shift, 2, 3))
    out = E.conv_layer(a.to(DEV), wt, "flowstem", 2, 3, 1, shift=shift, act=2, offset=b.to(DEV), engine=engine).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() < _tol(ref)

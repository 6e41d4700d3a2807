"""Checks at sizes the small-clip parity tests do not reach (-m gpu).

* 384x640 / 256x384 / 384x256 (not powers of two; partial tiles in every layer): key + cur graphs of Accel-18, -50,
  -101 and DFF against the CPU oracle.
* BASELINE.json's full 1024x2048 (graph-level ORACLE parity at this size is tests/test_gpu_fullsize_oracle.py):
  size-independent properties --
  the label map is bit-exactly the lowest-index argmax of the emitted score volume, production mode (no score
  volume) emits the same labels, two runs are bit-identical (deterministic split-K, fixed graphs), interval 1 equals
  the key graph on every frame, the key-frame lookahead equals the sequential loop, and warping with a zero / an
  integer flow field is the identity / a pure shift (and bit-exact against the oracle's sampler)."""
import numpy as np
import pytest
import torch

from accel_b200 import engine as E
from accel_b200 import scheduler, synthetic
from accel_b200.engine import Engine
from oracle import nets, ops
from parity_util import label_report

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-3


@pytest.mark.parametrize("version,H,W", [("18", 384, 640), ("50", 256, 384), ("101", 256, 384), ("dff", 384, 256)])
def test_odd_size_parity(version, H, W):
    params = synthetic.make_params(version)
    frames = synthetic.make_frames(2, H, W, stream=7)
    eng = Engine(version, H, W, params=params)
    dev = eng.torch_device
    d0, d1 = frames[0].to(dev), frames[1].to(dev)
    feat, feat2 = torch.empty(eng.feat_shape, device=dev), torch.empty(eng.feat_shape, device=dev)
    score = torch.empty(1, 19, H, W, device=dev)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        rk = nets.key_forward(params, frames[0])
        rc = nets.cur_forward(params, version, frames[1], frames[0], rk["res5c_relu_output"])
    eng.key_forward(d0, feat, score, label)
    assert (feat.cpu() - rk["res5c_relu_output"]).abs().max().item() < SCORE_TOL
    assert (score.cpu() - rk["croped_score_output"]).abs().max().item() < SCORE_TOL
    eng.cur_forward(d1, d0, rk["res5c_relu_output"].to(dev), feat2, score, label)
    assert torch.equal(feat2.cpu(), rc["warping_feat_output"]) or \
        (feat2.cpu() - rc["warping_feat_output"]).abs().max().item() < SCORE_TOL
    ref = rc[nets.output_key(version)]
    label_report(label, score.cpu(), ref, min_decided=0.995, max_mismatch_frac=5e-4)
    eng.close()


@pytest.fixture(scope="module")
def full():
    H, W = 1024, 2048
    eng = Engine("dff", H, W, params=synthetic.make_params("dff"))
    frames = [f.to(eng.torch_device) for f in synthetic.make_frames(4, H, W, stream=8)]
    yield eng, frames, H, W
    eng.close()


def test_full_size_labels_are_argmax_of_scores_and_deterministic(full):
    eng, frames, H, W = full
    dev = eng.torch_device
    feat = [torch.empty(eng.feat_shape, device=dev) for _ in range(3)]
    score = torch.empty(1, 19, H, W, device=dev)
    lab, lab2, lab3 = (torch.empty(H, W, dtype=torch.uint8, device=dev) for _ in range(3))
    eng.key_forward(frames[0], feat[0], score, lab)
    assert torch.equal(lab, score.argmax(dim=1)[0].to(torch.uint8))          # ties -> lowest index, like torch.argmax
    eng.key_forward(frames[0], feat[1], None, lab2)                           # production mode
    assert torch.equal(lab, lab2) and torch.equal(feat[0], feat[1])           # and bit-identical across runs
    eng.cur_forward(frames[1], frames[0], feat[0], feat[2], score, lab)
    assert torch.equal(lab, score.argmax(dim=1)[0].to(torch.uint8))
    eng.cur_forward(frames[1], frames[0], feat[0], feat[1], None, lab3)
    assert torch.equal(lab, lab3) and torch.equal(feat[2], feat[1])
    assert torch.isfinite(score).all() and torch.isfinite(feat[2]).all()


def test_full_size_interval_one_and_lookahead(full):
    eng, frames, H, W = full
    dev = eng.torch_device
    lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
    ref = []
    st = scheduler.StreamState(eng)
    for f in frames:                                                           # interval 2, sequential
        scheduler.segment_frame(eng, st, f, 2, "chained", lab)
        ref.append(lab.clone())
    st = scheduler.StreamState(eng)
    for i, f in enumerate(frames):                                             # same with the key-frame lookahead
        nk = frames[i + 2] if (i % 2 == 0 and i + 2 < len(frames)) else None
        scheduler.segment_frame(eng, st, f, 2, "chained", lab, next_key_data=nk)
        assert torch.equal(lab, ref[i]), "frame %d" % i
    feat = torch.empty(eng.feat_shape, device=dev)
    klab = torch.empty_like(lab)
    st = scheduler.StreamState(eng)
    for f in frames[:2]:                                                       # interval 1 == key graph every frame
        scheduler.segment_frame(eng, st, f, 1, "chained", lab)
        eng.key_forward(f, feat, None, klab)
        assert torch.equal(lab, klab)


def test_full_size_warp_identity_and_shift():
    c, h, w = 2048, 64, 128
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(1, c, h, w, generator=g).cuda()
    zero = torch.zeros(1, 2, h, w, device="cuda")
    # the reference's normalise / de-normalise round trip is not exact in fp32: identity up to a few ulps of blending
    assert torch.allclose(E.warp(feat, zero), feat, atol=1e-4)
    flow = zero.clone()
    flow[:, 0] = 5.0
    flow[:, 1] = -2.0
    out = E.warp(feat, flow)
    assert torch.allclose(out[:, :, 2:, : w - 5], feat[:, :, : h - 2, 5:], atol=1e-3)
    assert out[:, :, :1].abs().max().item() < 1e-3 and out[:, :, :, w - 4:].abs().max().item() < 1e-3   # zeros outside the map
    ref = ops.bilinear_sampler(feat.cpu(), ops.grid_generator_warp(flow.cpu()))
    assert torch.equal(out.cpu(), ref)                                   # and bit-exact against the oracle, at full size

"""Graph-level oracle parity at BASELINE.json's own size, 1024x2048 (-m gpu): one key frame and one cur frame of DFF,
Accel-18 and Accel-101 through the C ABI against the CPU oracle on the same synthetic frames -- the tile / split-K /
epilogue choices of the tcgen05 plans depend on the layer shapes, so the configurations the benchmark runs are the
ones compared here.  Score volume <= 1e-3 max-abs, features <= 1e-3, labels per tests/parity_util.py."""
import pytest
import torch

from accel_b200 import synthetic
from accel_b200.engine import Engine
from oracle import nets
from parity_util import SCORE_TOL, critical_mask, dcn_trace, label_report

pytestmark = pytest.mark.gpu
H, W = 1024, 2048


@pytest.fixture(scope="module")
def frames():
    return synthetic.make_frames(2, H, W, stream=5)


@pytest.fixture(scope="module")
def key_ref(frames):
    trace = []
    with torch.no_grad(), dcn_trace(trace):
        out = nets.key_forward(synthetic.make_params("dff"), frames[0])      # the key net is shared by all versions
    out["critical"] = critical_mask(trace, H, W)       # footprint of border-critical deformable samples (parity_util)
    return out


def _cur_ref(params, version, frames, key_ref):
    trace = []
    with torch.no_grad(), dcn_trace(trace):
        rc = nets.cur_forward(params, version, frames[1], frames[0], key_ref["res5c_relu_output"])
    return rc, critical_mask(trace, H, W)


@pytest.mark.parametrize("version", ["dff", "18", "101"])
def test_full_size_key_and_cur_against_oracle(version, frames, key_ref):
    params = synthetic.make_params(version)
    eng = Engine(version, H, W, params=params)
    dev = eng.torch_device
    d0, d1 = frames[0].to(dev), frames[1].to(dev)
    feat, feat2 = torch.empty(eng.feat_shape, device=dev), torch.empty(eng.feat_shape, device=dev)
    score = torch.empty(1, 19, H, W, device=dev)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    eng.key_forward(d0, feat, score, label)
    if not key_ref["critical"].any():
        assert (feat.cpu() - key_ref["res5c_relu_output"]).abs().max().item() < SCORE_TOL
    rep_k = label_report(label, score.cpu(), key_ref["croped_score_output"], exclude=key_ref["critical"])
    # cur frame on the GPU's OWN key feature, as the chained loop runs it; the oracle on its own
    eng.cur_forward(d1, d0, feat, feat2, score, label)
    rc, crit = _cur_ref(params, version, frames, key_ref)
    if not key_ref["critical"].any():
        assert (feat2.cpu() - rc["warping_feat_output"]).abs().max().item() < SCORE_TOL
    rep_c = label_report(label, score.cpu(), rc[nets.output_key(version)], exclude=crit | key_ref["critical"])
    print("full-size parity %s: key %r cur %r" % (version, rep_k, rep_c))
    eng.close()


@pytest.mark.parametrize("version", ["101", "18"])
def test_full_size_interval_plan_against_oracle(version, frames, key_ref):
    """The whole-interval plan (batched launches: other tile counts, other accumulation-chain choices than the
    frame-by-frame plans) against the oracle's chained schedule at 1024x2048: key frame + first cur frame."""
    params = synthetic.make_params(version)
    eng = Engine(version, H, W, params=params, interval=2)
    dev = eng.torch_device
    dfr = [f.to(dev) for f in frames]
    labels = torch.empty(2, H, W, dtype=torch.uint8, device=dev)
    scores = [torch.empty(1, 19, H, W, device=dev) for _ in range(2)]
    eng.interval_forward(dfr, labels, scores)
    rep_k = label_report(labels[0], scores[0].cpu(), key_ref["croped_score_output"], exclude=key_ref["critical"])
    rc, crit = _cur_ref(params, version, frames, key_ref)
    rep_c = label_report(labels[1], scores[1].cpu(), rc[nets.output_key(version)], exclude=crit | key_ref["critical"])
    print("full-size parity, interval plan %s: key %r cur %r" % (version, rep_k, rep_c))
    eng.close()

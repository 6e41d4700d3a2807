"""End-to-end parity of the key / cur plans and both keyframe schedules against the CPU oracle,
through the reference-shaped Predictor interface and the C ABI (-m gpu)."""
import numpy as np
import pytest
import torch

from accel_b200 import predictor as P
from accel_b200 import scheduler, synthetic
from accel_b200.engine import Engine
from oracle import nets, ops
from parity_util import label_report
from oracle import schedule as oracle_schedule

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-3          # north-star: max-abs on the fp32 score volume
H, W = 128, 256


def _label_check(label, ref_score, gpu_score):
    """tests/parity_util.py: zero mismatches wherever the oracle's top-2 margin exceeds twice the MEASURED score
    error, that set covers >= 99 % of these small frames, and the total number of flipped pixels is bounded."""
    return label_report(label, gpu_score, ref_score, min_decided=0.99, max_mismatch_frac=1e-3)


@pytest.fixture(scope="module")
def frames():
    return synthetic.make_frames(6, H, W)


@pytest.mark.parametrize("version", ["dff", "18", "34", "50", "101"])
def test_key_and_cur_graph_parity(version, frames):
    params = synthetic.make_params(version)
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
    _label_check(label.cpu().numpy(), rk["croped_score_output"], score.cpu())
    flow = eng.flownet(d1, d0)
    assert (flow.cpu() - rc["flow"]).abs().max().item() < 1e-4
    # feed the ORACLE's key feature so the cur graph is checked on identical inputs
    eng.cur_forward(d1, d0, rk["res5c_relu_output"].to(dev), feat2, score, label)
    assert (feat2.cpu() - rc["warping_feat_output"]).abs().max().item() < SCORE_TOL
    ref_score = rc[nets.output_key(version)]
    assert (score.cpu() - ref_score).abs().max().item() < SCORE_TOL
    _label_check(label.cpu().numpy(), ref_score, score.cpu())
    # production mode (no score volume, no feature copy) yields the same labels
    label2 = torch.empty_like(label)
    eng.cur_forward(d1, d0, rk["res5c_relu_output"].to(dev), None, None, label2)
    assert torch.equal(label, label2)
    eng.close()


@pytest.mark.parametrize("schedule", ["chained", "unchained"])
def test_reference_loop_through_predictor_shim(schedule, frames):
    version, interval = "18", 3
    params = synthetic.make_params(version)
    arg = dict(params)
    sym = P.accel_18()
    data_names = ["data", "data_key", "feat_key"]
    provide = [[("data", (1, 3, H, W)), ("data_key", (1, 3, H, W)), ("feat_key", (1, 2048, 1, 1))]]
    kp = P.Predictor(sym.get_key_test_symbol(None), data_names, [], context=[P.gpu(0)], provide_data=provide,
                     arg_params=arg, aux_params={})
    cp = P.Predictor(sym.get_cur_test_symbol(None), data_names, [], context=[P.gpu(0)], provide_data=provide,
                     arg_params=arg, aux_params={})
    assert kp.engine is cp.engine
    dev = kp.engine.torch_device
    got = scheduler.run_reference_loop(kp, cp, [f.to(dev) for f in frames], interval, version, schedule)
    with torch.no_grad():
        ref = oracle_schedule.run(params, version, frames, interval, schedule, keep=("label", "score", "feat"))
    for g, r in zip(got, ref):
        assert g["is_key"] == r["is_key"]
        assert (g["score"].cpu() - r["score"]).abs().max().item() < SCORE_TOL
        assert (g["feat"].cpu() - r["feat"]).abs().max().item() < SCORE_TOL
        _label_check(g["label"].cpu().numpy(), r["score"], g["score"].cpu())
        assert torch.equal(g["label"], g["label_output"])      # fused argmax == argmax of the emitted volume


def test_interval_one_is_key_graph_every_frame(frames):
    params = synthetic.make_params("dff")
    eng = Engine("dff", H, W, params=params)
    dev = eng.torch_device
    state = scheduler.StreamState(eng)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    ref_label = torch.empty_like(label)
    for f in frames[:3]:
        assert scheduler.segment_frame(eng, state, f.to(dev), 1, "chained", label)
        eng.key_forward(f.to(dev), None, None, ref_label)
        assert torch.equal(label, ref_label)
    eng.close()


def test_accel101_identity_correction_equals_r_branch(frames):
    # accel_101.py:276-289 initialises corr_weight = [0 | I], corr_bias = 0: the fused feature is then the
    # current frame's own R101 feature, so the cur graph must reproduce the key graph on that frame.
    params = synthetic.make_params("101")
    w = torch.zeros_like(params["corr_weight"])
    idx = torch.arange(2048)
    w[idx, 2048 + idx, 0, 0] = 1.0
    params["corr_weight"], params["corr_bias"] = w, torch.zeros(2048)
    eng = Engine("101", H, W, params=params)
    dev = eng.torch_device
    d0, d1 = frames[0].to(dev), frames[1].to(dev)
    feat = torch.empty(eng.feat_shape, device=dev)
    s_key, s_cur = torch.empty(1, 19, H, W, device=dev), torch.empty(1, 19, H, W, device=dev)
    eng.key_forward(d0, feat, None, None)
    eng.key_forward(d1, None, s_key, None)
    eng.cur_forward(d1, d0, feat, None, s_cur, None)
    assert (s_key - s_cur).abs().max().item() < 1e-4
    eng.close()


def test_errors_are_reported_not_swallowed(frames):
    params = synthetic.make_params("dff")
    eng = Engine("dff", H, W, params=params)
    dev = eng.torch_device
    with pytest.raises(ValueError):
        eng.key_forward(torch.zeros(1, 3, 64, 64, device=dev))
    with pytest.raises(TypeError):
        eng.key_forward(torch.zeros(1, 3, H, W))
    feat = torch.zeros(eng.feat_shape, device=dev)
    with pytest.raises(RuntimeError, match="alias"):
        eng.cur_forward(frames[0].to(dev), frames[0].to(dev), feat, feat, None, None)
    missing = dict(params)
    del missing["fc6_bias"]
    with pytest.raises(RuntimeError, match="fc6_bias"):
        Engine("dff", H, W, params=missing)
    eng.close()


@pytest.mark.parametrize("version", ["18", "34"])
def test_r_head_fold_on_and_off_agree_with_oracle(version, frames, monkeypatch):
    """`<v>_fc6 o <v>_feat_upsampling` runs as one composed 512 -> 1024 transposed conv by default (SURVEY.md 7-iii,
    accel_18.py:204-213); ACCEL_FOLD_FC6=0 runs the two layers as written.  Both must sit inside the score tolerance
    of the oracle (which always evaluates the graph as written) and agree with each other far inside it."""
    params = synthetic.make_params(version)
    with torch.no_grad():
        rk = nets.key_forward(params, frames[0])
        rc = nets.cur_forward(params, version, frames[1], frames[0], rk["res5c_relu_output"])
    ref_score = rc[nets.output_key(version)]
    scores = {}
    for fold in ("1", "0"):
        monkeypatch.setenv("ACCEL_FOLD_FC6", fold)
        eng = Engine(version, H, W, params=params)
        dev = eng.torch_device
        score = torch.empty(1, 19, H, W, device=dev)
        label = torch.empty(H, W, dtype=torch.uint8, device=dev)
        eng.cur_forward(frames[1].to(dev), frames[0].to(dev), rk["res5c_relu_output"].to(dev), None, score, label)
        assert (score.cpu() - ref_score).abs().max().item() < SCORE_TOL
        _label_check(label.cpu().numpy(), ref_score, score.cpu())
        scores[fold] = score.cpu()
        eng.close()
    assert (scores["1"] - scores["0"]).abs().max().item() < 2e-4

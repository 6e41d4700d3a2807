"""The commuted L head (accel_key_forward_lin / accel_cur_forward_lin, include/accel_b200.h) against the CPU oracle of
the reference graphs AS WRITTEN (accel_18.py:172-197: warp the 2048-channel feature, then fc6): the two orders of
the linear maps must agree inside the same 1e-3 score tolerance / decided-label rule as the plain plans (-m gpu)."""
import pytest
import torch
import torch.nn.functional as F

from accel_b200 import scheduler, synthetic
from accel_b200.engine import Engine
from oracle import nets, ops
from parity_util import label_report
from oracle import schedule as oracle_schedule

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-3
H, W = 128, 256


def _label_check(label, ref_score, gpu_score):
    """tests/parity_util.py: zero mismatches wherever the oracle's top-2 margin exceeds twice the MEASURED score
    error, that set covers >= 99 % of these small frames, and the total number of flipped pixels is bounded."""
    return label_report(label, gpu_score, ref_score, min_decided=0.99, max_mismatch_frac=1e-3)


def _g_of(params, feat):
    """fc6's linear part on the oracle's feature: W_fc6 * F, no bias, no ReLU."""
    return F.conv2d(feat, params["fc6_weight"])


@pytest.fixture(scope="module")
def frames():
    return synthetic.make_frames(7, H, W)


@pytest.mark.parametrize("flags", [0, 1])          # 1 = ACCEL_FLAG_NO_TENSOR_CORES: the CUDA-core cross-check engine
@pytest.mark.parametrize("version", ["dff", "18", "34", "50"])
def test_key_and_cur_lin_parity(version, flags, frames):
    if flags and version not in ("dff", "18"):
        pytest.skip("CUDA-core engine: two versions are enough")
    params = synthetic.make_params(version)
    eng = Engine(version, H, W, params=params, flags=flags)
    dev = eng.torch_device
    d0, d1 = frames[0].to(dev), frames[1].to(dev)
    g0, g1 = torch.empty(eng.g_shape, device=dev), torch.empty(eng.g_shape, device=dev)
    feat = torch.empty(eng.feat_shape, device=dev)
    score = torch.empty(1, 19, H, W, device=dev)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        rk = nets.key_forward(params, frames[0])
        rc = nets.cur_forward(params, version, frames[1], frames[0], rk["res5c_relu_output"])
        gk_ref = _g_of(params, rk["res5c_relu_output"])
        gc_ref = _g_of(params, rc["warping_feat_output"])
    # key plan: every ordinary output is unchanged, g_out = W*F on top
    eng.key_forward(d0, feat, score, label, g_out=g0)
    assert (feat.cpu() - rk["res5c_relu_output"]).abs().max().item() < SCORE_TOL
    assert (score.cpu() - rk["croped_score_output"]).abs().max().item() < SCORE_TOL
    gscale = max(1.0, gk_ref.abs().max().item())
    assert (g0.cpu() - gk_ref).abs().max().item() < SCORE_TOL * gscale
    _label_check(label.cpu().numpy(), rk["croped_score_output"], score.cpu())
    # cur plan fed the ORACLE's G: warp(W*F) + b == W*warp(F) + b
    eng.cur_forward_lin(d1, d0, gk_ref.to(dev), g1, score, label)
    assert (g1.cpu() - gc_ref).abs().max().item() < SCORE_TOL * gscale
    ref_score = rc[nets.output_key(version)]
    assert (score.cpu() - ref_score).abs().max().item() < SCORE_TOL
    _label_check(label.cpu().numpy(), ref_score, score.cpu())
    # production mode (no score volume, no carried G) gives the same labels
    label2 = torch.empty_like(label)
    eng.cur_forward_lin(d1, d0, gk_ref.to(dev), None, None, label2)
    assert torch.equal(label, label2)
    eng.close()


@pytest.mark.parametrize("schedule", ["chained", "unchained"])
@pytest.mark.parametrize("version", ["dff", "18"])
def test_linear_head_schedule_equals_oracle_loop(version, schedule, frames):
    interval = 3
    params = synthetic.make_params(version)
    eng = Engine(version, H, W, params=params)
    dev = eng.torch_device
    state = scheduler.StreamState(eng, linear_head=True)
    assert state.linear_head and state.carry_shape == eng.g_shape
    with torch.no_grad():
        ref = oracle_schedule.run(params, version, frames, interval, schedule, keep=("label", "score"))
    score = torch.empty(1, 19, H, W, device=dev)
    label = torch.empty(H, W, dtype=torch.uint8, device=dev)
    dframes = [f.to(dev) for f in frames]
    for f, r in zip(dframes, ref):
        is_key = scheduler.segment_frame(eng, state, f, interval, schedule, label, score_out=score)
        assert is_key == r["is_key"]
        assert (score.cpu() - r["score"]).abs().max().item() < SCORE_TOL
        _label_check(label.cpu().numpy(), r["score"], score.cpu())
        assert torch.equal(label, torch.argmax(score, dim=1).to(torch.uint8)[0])
    eng.close()


def test_linear_head_lookahead_equals_sequential(frames):
    """Key-frame lookahead + commuted head: same label maps as the sequential commuted loop, bit for bit."""
    interval = 3
    params = synthetic.make_params("18")
    eng = Engine("18", H, W, params=params)
    dev = eng.torch_device
    dframes = [f.to(dev) for f in frames]
    out = {}
    for look in (False, True):
        state = scheduler.StreamState(eng, linear_head=True)
        labels = []
        for i, f in enumerate(dframes):
            label = torch.empty(H, W, dtype=torch.uint8, device=dev)
            nk = dframes[i + interval] if (look and i % interval == 0 and i + interval < len(dframes)) else None
            scheduler.segment_frame(eng, state, f, interval, "chained", label, next_key_data=nk)
            labels.append(label)
        torch.cuda.synchronize(dev)
        out[look] = labels
    for a, b in zip(out[False], out[True]):
        assert torch.equal(a, b)
    eng.close()


def test_linear_head_pipeline_matches_plain_pipeline_on_decided_pixels(frames):
    """VideoPipeline(linear_head=True) from host uint8 frames: labels equal the plain pipeline's wherever the oracle
    margin rule decides (both are within 1e-3 of the same scores)."""
    interval = 3
    params = synthetic.make_params("dff")
    eng = Engine("dff", H, W, params=params)
    u8 = [torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(i)).pin_memory()
          for i in range(6)]
    plain = scheduler.VideoPipeline(eng, interval, "chained").segment_video(u8)
    lin = scheduler.VideoPipeline(eng, interval, "chained", linear_head=True).segment_video(u8)
    total = differ = 0
    for a, b in zip(plain, lin):
        total += a.numel()
        differ += int((a != b).sum())
    assert differ <= total * 0.002, "commuted head flips more than near-tie pixels: %d / %d" % (differ, total)
    eng.close()


def test_lin_errors(frames):
    params = synthetic.make_params("101")
    eng = Engine("101", H, W, params=params)
    dev = eng.torch_device
    assert not eng.supports_linear_head
    assert not scheduler.StreamState(eng, linear_head=True).linear_head        # falls back to the plain plans
    g = torch.zeros(eng.g_shape, device=dev)
    with pytest.raises(RuntimeError, match="Accel-101"):
        eng.cur_forward_lin(frames[0].to(dev), frames[0].to(dev), g, None, None, None)
    eng.close()
    params = synthetic.make_params("dff")
    eng = Engine("dff", H, W, params=params)
    g = torch.zeros(eng.g_shape, device=dev)
    with pytest.raises(RuntimeError, match="alias"):
        eng.cur_forward_lin(frames[0].to(dev), frames[0].to(dev), g, g, None, None)
    with pytest.raises(ValueError):
        eng.cur_forward_lin(frames[0].to(dev), frames[0].to(dev), torch.zeros(eng.feat_shape, device=dev))
    eng.close()

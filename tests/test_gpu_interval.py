"""Whole-interval plan (accel_plan_interval / accel_interval_forward, -m gpu): one key interval issued as one launch
sequence with the per-frame chains running concurrently must give what the frame-by-frame loop gives -- against the
CPU oracle's chained schedule (score <= 1e-3, labels per tests/parity_util.py) and against the sequential CUDA loop."""
import pytest
import torch

from accel_b200 import scheduler, synthetic
from accel_b200.engine import Engine
from oracle import schedule as oracle_schedule
from parity_util import SCORE_TOL, label_report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("version,interval", [("dff", 3), ("18", 3), ("34", 2), ("50", 2), ("101", 3)])
def test_interval_plan_matches_oracle_and_sequential_loop(version, interval):
    H, W = 128, 256
    params = synthetic.make_params(version)
    frames = synthetic.make_frames(interval, H, W, stream=2)
    eng = Engine(version, H, W, params=params, interval=interval)
    dev = eng.torch_device
    dfr = [f.to(dev) for f in frames]
    labels = torch.empty(interval, H, W, dtype=torch.uint8, device=dev)
    scores = [torch.empty(1, 19, H, W, device=dev) for _ in range(interval)]
    st = scheduler.StreamState(eng)
    for rep in range(3):                                   # eager first use, capture, replay
        labels.zero_()
        scheduler.segment_interval(eng, st, dfr, labels, scores)
    with torch.no_grad():
        ref = oracle_schedule.run(params, version, frames, interval, "chained", keep=("label", "score"))
    for t in range(interval):
        rep = label_report(labels[t], scores[t].cpu(), ref[t]["score"], min_decided=0.99, max_mismatch_frac=1e-3)
        print("interval plan %s frame %d: %r" % (version, t, rep))
    # the sequential loop of the same engine: same labels wherever its own scores are decided
    st2 = scheduler.StreamState(eng)
    lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
    sc = torch.empty(1, 19, H, W, device=dev)
    for t in range(interval):
        scheduler.segment_frame(eng, st2, dfr[t], interval, "chained", lab, sc)
        assert (sc - scores[t]).abs().max().item() < SCORE_TOL
        top2 = sc.topk(2, dim=1).values
        decided = (top2[:, 0] - top2[:, 1])[0] > 2 * (sc - scores[t]).abs().max()
        assert torch.equal(lab[decided], labels[t][decided])
    # production mode (no score volumes) gives the same label maps, and is deterministic
    labels2 = torch.empty_like(labels)
    scheduler.segment_interval(eng, st, dfr, labels2)
    scheduler.segment_interval(eng, st, dfr, labels2)
    assert torch.equal(labels, labels2)
    eng.close()


def test_interval_plan_full_size_accel101_equals_sequential_loop():
    H, W, I = 1024, 2048, 5
    eng = Engine("101", H, W, params=synthetic.make_params("101"), interval=I)
    dev = eng.torch_device
    dfr = [f.to(dev) for f in synthetic.make_frames(I, H, W, stream=6)]
    labels = torch.empty(I, H, W, dtype=torch.uint8, device=dev)
    st = scheduler.StreamState(eng)
    for rep in range(3):
        scheduler.segment_interval(eng, st, dfr, labels)
    st2 = scheduler.StreamState(eng)
    lab = torch.empty(H, W, dtype=torch.uint8, device=dev)
    flips = 0
    for t in range(I):
        scheduler.segment_frame(eng, st2, dfr[t], I, "chained", lab)
        flips += int((lab != labels[t]).sum())
    print("interval plan vs frame-by-frame loop, Accel-101 1024x2048: %d of %d labels differ" % (flips, I * H * W))
    assert flips <= 2e-4 * I * H * W
    hits, misses = eng.graph_cache_stats()
    assert misses <= 4 and hits >= 1
    eng.close()


def test_interval_api_errors():
    eng = Engine("dff", 128, 256, params=synthetic.make_params("dff"))
    with pytest.raises(RuntimeError):
        eng.interval_forward([], [])
    with pytest.raises(RuntimeError):
        eng.plan_interval(3)                                   # after finalize
    eng.close()

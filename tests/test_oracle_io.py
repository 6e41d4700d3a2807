"""CPU checks of the ingest / metric restatements (oracle/io.py) against hand-computed values, and that the
synthetic generator's `transform` is the same function (so fixtures and the GPU ingest path agree)."""
import numpy as np
import torch

from accel_b200 import synthetic
from oracle import io as oio

MEANS = synthetic.PIXEL_MEANS_BGR


def test_transform_hand_values():
    im = np.zeros((2, 3, 3), dtype=np.uint8)
    im[0, 0] = (10, 20, 30)            # B, G, R
    im[1, 2] = (255, 0, 128)
    t = oio.transform(im, MEANS)
    assert t.shape == (1, 3, 2, 3) and t.dtype == np.float32
    assert t[0, 0, 0, 0] == np.float32(30 - 123.15)          # channel 0 = R - mean_R
    assert t[0, 1, 0, 0] == np.float32(20 - 115.90)
    assert t[0, 2, 0, 0] == np.float32(10 - 103.06)
    assert t[0, 0, 1, 2] == np.float32(128 - 123.15) and t[0, 2, 1, 2] == np.float32(255 - 103.06)
    assert t[0, 1, 0, 1] == np.float32(0 - 115.90)


def test_transform_is_float64_then_float32():
    # all 256 byte values per channel: float64 subtraction, one rounding (image.py:231-234, demo.py:185)
    v = np.arange(256, dtype=np.uint8)
    im = np.stack([v, v, v], axis=-1)[None]                  # (1,256,3)
    t = oio.transform(im, MEANS)
    for i in range(3):
        want = (v.astype(np.float64) - MEANS[2 - i]).astype(np.float32)
        assert np.array_equal(t[0, i, 0], want)


def test_synthetic_transform_matches_oracle():
    frames = synthetic.make_frames_u8(2, 64, 128, stream=3)
    for f in frames:
        a = synthetic.transform(f).numpy()
        b = oio.transform(f.numpy(), MEANS)
        assert np.array_equal(a, b)


def test_fast_hist_against_loops():
    rng = np.random.RandomState(0)
    n = 19
    pred = rng.randint(0, n, size=(37, 53)).astype(np.uint8)
    label = rng.randint(0, n, size=(37, 53)).astype(np.uint8)
    label[rng.rand(37, 53) < 0.2] = 255                      # Cityscapes ignore label
    h = oio.fast_hist(pred.flatten(), label.flatten(), n)
    want = np.zeros((n, n), dtype=np.int64)
    for p, l in zip(pred.flatten(), label.flatten()):
        if l < n:
            want[l, p] += 1
    assert np.array_equal(h, want)
    assert h.sum() == (label < n).sum()


def test_per_class_iu_and_miou():
    h = np.array([[3, 1, 0], [0, 2, 0], [0, 0, 0]], dtype=np.int64)
    iu = oio.per_class_iu(h)
    assert np.allclose(iu[:2], [3 / 4, 2 / 3]) and np.isnan(iu[2])
    assert oio.mean_iou(h) == round((3 / 4 + 2 / 3) / 2 * 100, 2)


def test_torch_confusion_helper_matches_fast_hist():
    from accel_b200 import scheduler
    g = torch.Generator().manual_seed(1)
    pred = torch.randint(0, 19, (64, 64), generator=g, dtype=torch.uint8)
    label = torch.randint(0, 19, (64, 64), generator=g, dtype=torch.uint8)
    label[::7] = 255
    got = scheduler.confusion_matrix(pred, label, 19).numpy()
    assert np.array_equal(got, oio.fast_hist(pred.numpy().flatten(), label.numpy().flatten(), 19))


def test_resize_golden_vectors_are_what_cv2_produces():
    """tests/golden/resize_vectors.npz must stay what the installed cv2 computes (the GPU test compares against the file)."""
    import os

    import numpy as np
    import pytest
    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_vectors.npz"))
    for i in range(int(g["n"])):
        fx = float(g["fx_%d" % i])
        assert np.array_equal(cv2.resize(g["src_%d" % i], None, None, fx=fx, fy=fx, interpolation=cv2.INTER_LINEAR), g["dst_%d" % i])

"""Palettised PNG writer (accel_b200/visualize.py) against PIL's reader and the Cityscapes colour table."""
import numpy as np
import pytest

from accel_b200 import visualize


def test_palette_table():
    pal = visualize.getpallete(256).reshape(256, 3)
    assert tuple(pal[0]) == (128, 64, 128) and tuple(pal[13]) == (0, 0, 142) and tuple(pal[18]) == (119, 11, 32)
    assert not pal[19:].any() and pal.dtype == np.uint8
    assert visualize.getpallete(5).shape == (15,)


def test_png_round_trip(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(0)
    pred = rng.randint(0, 19, (37, 53)).astype(np.uint8)
    path = visualize.save_segmentation(pred, str(tmp_path / "seg.png"))
    im = Image.open(path)
    assert im.mode == "P" and im.size == (53, 37)
    assert np.array_equal(np.asarray(im), pred)
    assert list(im.getpalette()[:57]) == list(visualize.getpallete(19))
    rgb = np.asarray(im.convert("RGB"))
    assert tuple(rgb[0, 0]) == tuple(visualize.getpallete(19).reshape(19, 3)[pred[0, 0]])


def test_rejects_non_label_maps(tmp_path):
    with pytest.raises(TypeError):
        visualize.save_segmentation(np.zeros((4, 4), dtype=np.float32), str(tmp_path / "x.png"))

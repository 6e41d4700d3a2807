"""Host logic of accel_b200/loader.py that needs no GPU: config defaults and the per-GPU result merge."""
import numpy as np

from accel_b200 import loader
from oracle import io as oio


def test_default_config_fields():
    cfg = loader.default_config(key_frame_interval=5)
    assert cfg.TEST.KEY_FRAME_INTERVAL == 5 and cfg.SCALES[0] == (1024, 2048)
    assert list(cfg.network.PIXEL_MEANS) == [103.06, 115.90, 123.15] and cfg.network.DFF_FEAT_DIM == 2048
    assert cfg.dataset.NUM_CLASSES == 19 and cfg.network.IMAGE_STRIDE == 0


def test_merge_results_matches_oracle_miou():
    rng = np.random.RandomState(0)
    hs = [rng.randint(0, 50, (19, 19)).astype(np.int64) for _ in range(3)]
    res = [{"hist": h, "frame_ids": np.arange(i * 4, i * 4 + 4)} for i, h in enumerate(hs)]
    m = loader.merge_results(res)
    total = hs[0] + hs[1] + hs[2]
    assert np.array_equal(m["hist"], total) and list(m["frame_ids"]) == list(range(12))
    assert m["mIoU"] == oio.mean_iou(total)
    assert np.allclose(m["ious"], oio.per_class_iu(total) * 100, equal_nan=True)

"""Frame ingest and accuracy bookkeeping either side of the per-frame graphs, restated in numpy exactly as
the reference writes them (they are plain numpy there too, so these are line-for-line semantics, not
[MXNet-ext] assumptions).  Test infrastructure only.

transform      -- lib/utils/image.py:224-235 (called from dff_deeplab/demo.py:175); the float64 result
                  becomes float32 at mx.nd.array (demo.py:185).
fast_hist      -- dff_deeplab/demo.py:50-53.
per_class_iu   -- dff_deeplab/demo.py:55-56.
"""
from __future__ import annotations

import numpy as np


def transform(im, pixel_means):
    """im: (H,W,3) BGR, any numeric dtype; pixel_means: [B,G,R].  Returns (1,3,H,W) float32 RGB minus mean
    (float64 arithmetic, one rounding to float32 -- image.py:231-234 then demo.py:185)."""
    im = np.asarray(im)
    im_tensor = np.zeros((1, 3, im.shape[0], im.shape[1]))
    for i in range(3):
        im_tensor[0, i, :, :] = im[:, :, 2 - i] - pixel_means[2 - i]
    return im_tensor.astype(np.float32)


def fast_hist(pred, label, n):
    """Confusion counts (rows = label, columns = prediction) over the pixels whose label is a class id."""
    pred = np.asarray(pred).reshape(-1)
    label = np.asarray(label).reshape(-1)
    k = (label >= 0) & (label < n)
    return np.bincount(n * label[k].astype(int) + pred[k], minlength=n ** 2).reshape(n, n)


def per_class_iu(hist):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.true_divide(np.diag(hist), (hist.sum(1) + hist.sum(0) - np.diag(hist)))


def mean_iou(hist):
    """`round(np.nanmean(per_class_iu(hist)) * 100, 2)` as printed at demo.py:274-281."""
    return round(float(np.nanmean(per_class_iu(hist))) * 100, 2)

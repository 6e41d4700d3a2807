"""Palettised PNG output of a label map: the three lines after the argmax in dff_deeplab/demo.py:252-256
(`Image.fromarray(pred)`, `putpalette(getpallete(256))`, `save`).  Host side, numpy + zlib only (PIL is not
required to write; the tests read the files back with it)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

# Colour of Cityscapes train id 0..18: what demo.py:58-104 computes as pallete_raw[train2regular[i] + 1].
_TRAIN_ID_COLOURS = (
    (128, 64, 128), (244, 35, 232), (70, 70, 70), (102, 102, 156), (190, 153, 153), (153, 153, 153), (250, 170, 30),
    (220, 220, 0), (107, 142, 35), (152, 251, 152), (70, 130, 180), (220, 20, 60), (255, 0, 0), (0, 0, 142), (0, 0, 70),
    (0, 60, 100), (0, 80, 100), (0, 0, 230), (119, 11, 32))


def getpallete(num_cls):
    """Flat uint8 palette of length 3 * num_cls; entries past the 19 train ids are black (demo.py:58-104)."""
    pal = np.zeros((num_cls, 3), dtype=np.uint8)
    n = min(num_cls, len(_TRAIN_ID_COLOURS))
    pal[:n] = np.asarray(_TRAIN_ID_COLOURS[:n], dtype=np.uint8)
    return pal.reshape(-1)


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def save_segmentation(pred, path, num_cls=256, level=6):
    """Writes `pred` (H, W) uint8 label map as an 8-bit palettised PNG with getpallete(num_cls) -- what
    `segmentation_result.save(output_dir + '/seg_' + im_filename)` produces (demo.py:253-256)."""
    pred = np.ascontiguousarray(np.asarray(pred))
    if pred.dtype != np.uint8 or pred.ndim != 2:
        raise TypeError("pred must be a (H, W) uint8 array (np.uint8(np.squeeze(argmax)), demo.py:252)")
    h, w = pred.shape
    raw = np.empty((h, w + 1), dtype=np.uint8)
    raw[:, 0] = 0                                      # filter type 0 on every scanline
    raw[:, 1:] = pred
    png = b"\x89PNG\r\n\x1a\n"
    png += _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 3, 0, 0, 0))       # 8 bit, colour type 3 = palette
    png += _chunk(b"PLTE", getpallete(num_cls).tobytes())
    png += _chunk(b"IDAT", zlib.compress(raw.tobytes(), level))
    png += _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)
    return path

#!/usr/bin/env python
"""Golden vectors for accel_resize_bgr: outputs of the REAL cv2.resize(..., interpolation=cv2.INTER_LINEAR) -- the call
lib/utils/image.py:211 makes -- on seeded random uint8 BGR images, written to tests/golden/resize_vectors.npz.

    python tests/golden/make_resize_vectors.py        # needs cv2 (4.13.0 in the build container)

Cases cover up- and down-scaling, non-integer scales, the exact 2x decimation OpenCV computes as INTER_AREA, scales
below 0.5, and the reference's own scale rule (short side -> 1024 unless the long side would pass 2048)."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (height, width, fx)
    (37, 53, 1.7), (48, 64, 2.0), (60, 80, 1.28), (54, 96, 0.9481481481481482), (33, 51, 1.0 / 1.3), (64, 96, 0.5),
    (70, 90, 0.3), (5, 7, 3.3), (41, 59, 1.0), (2, 2, 4.0),
]


def main():
    rng = np.random.default_rng(20261017)
    out = {"cv2_version": np.array(cv2.__version__), "n": np.array(len(CASES))}
    for i, (h, w, fx) in enumerate(CASES):
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        dst = cv2.resize(im, None, None, fx=fx, fy=fx, interpolation=cv2.INTER_LINEAR)
        out["src_%d" % i], out["fx_%d" % i], out["dst_%d" % i] = im, np.array(fx), dst
    np.savez_compressed(os.path.join(HERE, "resize_vectors.npz"), **out)
    print("wrote %d cases with cv2 %s" % (len(CASES), cv2.__version__))


if __name__ == "__main__":
    main()

"""Golden vectors for the ROI crop: cv2.warpAffine outputs (the reference's crop_resize_by_warp_affine,
core/utils/data_utils.py:115-133) on a small seeded image -> tests/golden/crop_golden.npz."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.RandomState(7)
img = rng.randint(0, 256, (96, 128, 3)).astype(np.uint8)
Ms, outs, crops = [], [], {}
for i in range(6):
    cx, cy, scale = rng.uniform(-10, 138), rng.uniform(-10, 106), float(rng.uniform(12, 160))
    out = 64 if i % 2 else 32
    s = out / scale
    M = np.array([[s, 0, out * 0.5 - cx * s], [0, s, out * 0.5 - cy * s]], np.float64)
    if i == 4:
        a = 0.4
        M[:, :2] = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) * s
    Ms.append(M)
    outs.append(out)
    crops["crop_%d" % i] = cv2.warpAffine(img, M, (out, out), flags=cv2.INTER_LINEAR)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "crop_golden.npz"), img=img, M=np.stack(Ms), out=np.array(outs), **crops)
print("wrote crop_golden.npz", cv2.__version__)

"""Generate tests/golden/fps_golden.npz with the REFERENCE's own FPS implementation
(core/csrc/fps/src/farthest_point_sampling.cpp compiled into oracle/_ref/libfps_ref.so by oracle/build_ref.py,
flags of core/csrc/fps/setup.py:5-7).  init_center entry point only (the other one is time-seeded)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfps_ref.so"))
out = {}
rs = np.random.RandomState(123)
cases = [(64, 8, 1.0), (1000, 32, 0.2), (4096, 64, 0.2), (8192, 64, 0.2), (333, 333, 5.0)]
for i, (pn, sn, scale) in enumerate(cases):
    pts = ((rs.rand(pn, 3) - 0.5) * scale).astype(np.float32)
    if i == 1:
        pts[100:150] = pts[7]  # duplicates
    idx = np.zeros(sn, np.int32)
    ref.farthest_point_sampling_init_center(pts.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p), pn, sn)
    out[f"pts_{i}"] = pts
    out[f"idx_{i}"] = idx
out["n_cases"] = np.int64(len(cases))
path = os.path.join(ROOT, "tests", "golden", "fps_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))

"""ROI pre-processing: batched GPU crops vs the reference's per-ROI cv2.warpAffine host loop (predictor_gdrn.py:417-438)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import native_ops as NO
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
H, W, n = 480, 640, 64
image = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
depth = rng.rand(H, W).astype(np.float32)
xx, yy = np.meshgrid(np.linspace(0, 1, W, endpoint=False, dtype=np.float32), np.linspace(0, 1, H, endpoint=False, dtype=np.float32))
coord = np.stack([xx, yy], 2)
cs = np.stack([rng.uniform(60, W - 60, n), rng.uniform(60, H - 60, n)], 1)
scales = rng.uniform(60, 300, n)
M256 = np.stack([NO.get_affine_transform(c, float(s), 0, 256) for c, s in zip(cs, scales)])
M64 = np.stack([NO.get_affine_transform(c, float(s), 0, 64) for c, s in zip(cs, scales)])
img_d, dep_d, crd_d = torch.from_numpy(image).to(dev), torch.from_numpy(depth).to(dev), torch.from_numpy(coord).to(dev)
def gpu():
    NO.crop_resize_image(img_d, M256, 256); NO.crop_resize_float(crd_d, M64, 64); NO.crop_resize_float(dep_d, M256, 256, nearest=True)
for _ in range(5): gpu()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): gpu()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
out_bytes = n * (3 * 256 * 256 + 2 * 64 * 64 + 256 * 256) * 4
res = {"n_rois": n, "gpu_ms_per_64_rois": ms, "gpu_rois_per_s": n / ms * 1e3, "output_bytes": out_bytes,
       "achieved_GBps_written": out_bytes / ms / 1e6, "note": "3 launches incl. host-side M upload (64 x 48 B) each"}
try:
    import cv2
    t = time.perf_counter()
    for i in range(n):
        a = cv2.warpAffine(image, M256[i], (256, 256), flags=cv2.INTER_LINEAR).transpose(2, 0, 1)
        a = ((a - np.zeros((3, 1, 1))) / np.full((3, 1, 1), 255.0)).astype("float32")
        cv2.warpAffine(coord, M64[i], (64, 64), flags=cv2.INTER_LINEAR); cv2.warpAffine(depth, M256[i], (256, 256), flags=cv2.INTER_NEAREST)
    dt = time.perf_counter() - t
    res["cv2_host_loop_ms_per_64_rois"] = dt * 1e3
    res["cv2_rois_per_s"] = n / dt
except ImportError:
    pass
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "crop_bench.json"), "w"), indent=1)

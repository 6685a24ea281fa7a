"""Stand-in for BASELINE.md's R-GPU-5 / R-GPU-64 rows: the reference forward as plain eager PyTorch on the GPU.

The reference model cannot be imported (timm/mmcv/detectron2 absent), so this times the torch restatement from
oracle/ (same nn.functional calls the reference makes: cuDNN convs, cuBLAS linears, native LN/GN/GELU) in fp32 with
torch's default TF32-for-convs setting, exactly like core/gdrn_modeling/engine/gdrn_evaluator.py:707-751 times it
(perf_counter + cuda.synchronize).  A reported baseline for DESIGN.md -- NOT part of the product path or bench.py."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200.synthetic import make_batch, make_state_dict  # noqa: E402
from oracle import gdrn_model_oracle as O  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    sd = {k: v.to(dev) for k, v in make_state_dict().items()}
    out = {}
    for B in (5, 64):
        batch = {k: v.to(dev) for k, v in make_batch(B=B, seed=1).items()}

        def fwd():
            feat = O.convnext_features(sd, batch["roi_img"])
            vis, full, cx, cy, cz, region = O.geo_head(sd, feat)
            vis, full, cx, cy, cz, region = O.class_gather(vis, full, cx, cy, cz, region, batch["roi_classes"].cpu())
            coor = torch.cat([cx, cy, cz, batch["roi_coord_2d"]], dim=1)
            rs = torch.softmax(region[:, 1:], dim=1)
            rot6, t_ = O.conv_pnp_net(sd, coor, rs, batch["roi_extents"])
            return O.rot6d_to_mat_batch(rot6), t_

        with torch.no_grad():
            for _ in range(5):
                fwd()
            torch.cuda.synchronize()
            n = 20
            t0 = time.perf_counter()
            for _ in range(n):
                fwd()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
        out[f"bs{B}"] = {"ms_per_batch": dt * 1e3, "rois_per_s": B / dt}
    out["note"] = ("eager PyTorch fp32 (TF32 convs allowed = torch default) restatement of the reference forward on this "
                   "GPU, GPU part only (no allo->ego host loop); stand-in for the un-importable reference model")
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Scratch: dwconv phase trace for one forward at B=64 (GDRN_DW_TRACE=1)."""
import os, sys
os.environ.setdefault("GDRN_DW_TRACE", "1")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, default_cfg
from gdrnpp_bop2022_b200.synthetic import make_batch, make_state_dict
dev = torch.device("cuda:0")
m = GDRN_DoubleMask(default_cfg(), max_batch=64); m.load_state_dict(make_state_dict()); m.to(dev)
b = {k: v.to(dev) for k, v in make_batch(B=64, seed=0).items()}
for i in range(2):
    sys.stderr.write("=== forward %d\n" % i)
    m(b["roi_img"], roi_classes=b["roi_classes"], roi_coord_2d=b["roi_coord_2d"], roi_cams=b["roi_cams"], roi_centers=b["roi_centers"],
      roi_whs=b["roi_whs"], roi_extents=b["roi_extents"], resize_ratios=b["resize_ratios"])
    torch.cuda.synchronize()

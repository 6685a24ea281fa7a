"""Scratch: split-bf16 (bf16x3) precision mode vs the fp32 oracle + throughput of both modes."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, default_cfg  # noqa: E402
from gdrnpp_bop2022_b200.synthetic import make_batch, make_state_dict  # noqa: E402
from oracle import gdrn_model_oracle as O  # noqa: E402  (scratch checker)

dev = torch.device("cuda:0")
res = {}
sd = make_state_dict()
B = 8
batch = make_batch(B=B, seed=11)
torch.set_num_threads(max(1, os.cpu_count() or 1))
with torch.no_grad():
    ref = O.gdrn_forward(sd, batch, return_maps=True, return_intermediate=True)
gb = {k: v.to(dev) for k, v in batch.items()}
kw = dict(roi_classes=gb["roi_classes"], roi_coord_2d=gb["roi_coord_2d"], roi_cams=gb["roi_cams"],
          roi_centers=gb["roi_centers"], roi_whs=gb["roi_whs"], roi_extents=gb["roi_extents"],
          resize_ratios=gb["resize_ratios"])
for prec in ("bf16", "bf16x3"):
    try:
        m = GDRN_DoubleMask(default_cfg(with_maps=True), max_batch=64, precision=prec)
        m.load_state_dict(sd)
        m.to(dev)
        out = m(gb["roi_img"], return_raw=True, **kw)
        torch.cuda.synchronize()
        r = {}
        x3 = m.debug_read("stage3_x", B, B * 64 * 1024).reshape(B, 8, 8, 1024).permute(0, 3, 1, 2).cpu()
        r["stage3_rel"] = float((x3 - ref["conv_feat"]).norm() / ref["conv_feat"].norm())
        raw = out["raw"].cpu()
        r["rot6d_maxabs"] = float((raw[:, :6] - ref["rot6d"]).abs().max())
        r["t_maxabs"] = float((raw[:, 6:] - ref["t_"]).abs().max())
        for k in ("mask", "full_mask", "coor_x", "coor_y", "coor_z", "region"):
            r[k] = float((out[k].cpu() - ref[k]).abs().max())
        d = (out["rot"].cpu().double() - ref["rot"].double()).flatten(1).norm(dim=1)
        r["rot_err_rad"] = float((2 * torch.asin((d / (2 * 2 ** 0.5)).clamp(max=1.0))).max())
        r["trans_maxabs"] = float((out["trans"].cpu() - ref["trans"]).abs().max())
        # throughput at B=64
        b64 = {k: v.to(dev) for k, v in make_batch(B=64, seed=0).items()}
        kw64 = dict(roi_classes=b64["roi_classes"], roi_coord_2d=b64["roi_coord_2d"], roi_cams=b64["roi_cams"],
                    roi_centers=b64["roi_centers"], roi_whs=b64["roi_whs"], roi_extents=b64["roi_extents"],
                    resize_ratios=b64["resize_ratios"])
        for _ in range(3):
            m(b64["roi_img"], **kw64)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(b64["roi_img"], **kw64)
        e1.record()
        torch.cuda.synchronize()
        r["ms_per_batch64"] = e0.elapsed_time(e1) / 10
        r["rois_per_s"] = 64e3 / r["ms_per_batch64"]
        res[prec] = r
    except Exception as e:  # noqa: BLE001
        import traceback
        res[prec] = {"exception": repr(e), "tb": traceback.format_exc()[-1500:]}
    print(prec, json.dumps(res[prec]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "precise.json"), "w"), indent=1)

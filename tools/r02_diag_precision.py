"""Where does the parity-mode (bf16x3) rotation error come from?  Per-ROI diagnostics at B = 64 (bench batch, seed 0):
rotation error vs the fp32 CPU oracle, raw rot6d error, conditioning (|a1|, sin of the angle between a1 and a2), for the
weights of make_state_dict(); plus the same oracle graph run as eager fp32 PyTorch on the GPU (cuDNN / cuBLAS, TF32 off
and on) against the CPU oracle -- what 'the reference's own CUDA build' differs from its CPU path by."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, default_cfg  # noqa: E402
from gdrnpp_bop2022_b200.synthetic import make_batch, make_state_dict  # noqa: E402
from oracle import gdrn_model_oracle as O  # noqa: E402


def rot_err(Ra, Rb):
    d = (Ra.double() - Rb.double()).flatten(1).norm(dim=1)
    return 2 * torch.asin((d / (2 * 2 ** 0.5)).clamp(max=1.0))


def main():
    dev = torch.device("cuda:0")
    B = int(os.environ.get("DIAG_B", "64"))
    sd = make_state_dict()
    batch = make_batch(B=B, seed=0)
    with torch.no_grad():
        ref = O.gdrn_forward(sd, batch, return_intermediate=True)
    a1, a2 = ref["rot6d"][:, :3].double(), ref["rot6d"][:, 3:].double()
    n1 = a1.norm(dim=1)
    sin12 = torch.linalg.cross(a1 / n1[:, None], a2 / a2.norm(dim=1, keepdim=True)).norm(dim=1)
    out = {"B": B, "a1_norm_min": n1.min().item(), "a1_norm_median": n1.median().item(), "sin12_min": sin12.min().item()}
    gb = {k: v.to(dev) for k, v in batch.items()}
    for prec in ("bf16x3", "bf16"):
        m = GDRN_DoubleMask(default_cfg(), max_batch=B, precision=prec)
        m.load_state_dict(sd)
        m.to(dev)
        o = m(gb["roi_img"], roi_classes=gb["roi_classes"], roi_coord_2d=gb["roi_coord_2d"], roi_cams=gb["roi_cams"],
              roi_centers=gb["roi_centers"], roi_whs=gb["roi_whs"], roi_extents=gb["roi_extents"],
              resize_ratios=gb["resize_ratios"], return_raw=True)
        torch.cuda.synchronize()
        re = rot_err(o["rot"].cpu(), ref["rot"])
        raw = (o["raw"].cpu()[:, :6] - ref["rot6d"]).abs().max(dim=1)[0]
        worst = int(re.argmax())
        out[prec] = {"rot_err_max": re.max().item(), "rot_err_median": re.median().item(),
                     "rot_err_p90": re.kthvalue(max(1, int(0.9 * B))).values.item(),
                     "raw6d_err_max": raw.max().item(), "raw6d_err_median": raw.median().item(),
                     "t_err_max": (o["trans"].cpu() - ref["trans"]).abs().max().item(),
                     "worst_roi": {"idx": worst, "a1_norm": n1[worst].item(), "sin12": sin12[worst].item(), "raw_err": raw[worst].item()},
                     "n_over_1e-4": int((re > 1e-4).sum())}
        del m
    # eager fp32 torch on the GPU (the stand-in for the reference's CUDA build) vs the CPU oracle
    sdg = {k: v.to(dev) for k, v in sd.items()}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False   # torch default: TF32 for cuDNN convs only
        with torch.no_grad():
            feat = O.convnext_features(sdg, gb["roi_img"])
            vis, full, cx, cy, cz, region = O.geo_head(sdg, feat)
            vis, full, cx, cy, cz, region = O.class_gather(vis, full, cx, cy, cz, region, batch["roi_classes"])
            coor = torch.cat([cx, cy, cz, gb["roi_coord_2d"]], dim=1)
            rs = torch.softmax(region[:, 1:], dim=1)
            rot6, t_ = O.conv_pnp_net(sdg, coor, rs, gb["roi_extents"])
            Rm = O.rot6d_to_mat_batch(rot6.cpu())
        Rref = O.rot6d_to_mat_batch(ref["rot6d"])
        re = rot_err(Rm, Rref)
        out["eager_gpu_fp32_tf32conv_%s" % ("on" if tf32 else "off")] = {
            "rot_err_max_allo": re.max().item(), "rot_err_median_allo": re.median().item(),
            "raw6d_err_max": (rot6.cpu() - ref["rot6d"]).abs().max().item()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Scratch GPU bring-up script (not a test): GEMM kernel vs torch, model vs oracle, quick op sanity."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import _lib  # noqa: E402
from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, default_cfg  # noqa: E402
from gdrnpp_bop2022_b200.synthetic import make_batch, make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
results = {}
dev = torch.device("cuda:0")
L = _lib.lib()


def sync():
    torch.cuda.synchronize()


def gemm_case(M, N, K, block_n, epi, out_f32):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    gamma = torch.rand(N, generator=g).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev)
    ref = A.float() @ W.float().t() + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = resid + gamma * ref
    is_f32 = (epi == 2) or (epi == 0 and out_f32)
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if is_f32 else torch.bfloat16, device=dev)
    rc = L.gdrn_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(resid), _lib.ptr(out),
                          M, N, K, epi, int(out_f32), block_n, _lib.current_stream())
    if rc != 0:
        return {"rc": rc, "err": _lib.last_error()}
    sync()
    o = out.float()
    err = (o - ref).abs()
    return {"max_abs": float(err.max()), "mean_abs": float(err.mean()), "ref_absmax": float(ref.abs().max()),
            "nan": int(torch.isnan(o).sum())}


def run(name, fn):
    t = time.time()
    try:
        results[name] = fn()
    except Exception as e:  # noqa: BLE001
        results[name] = {"exception": repr(e), "tb": traceback.format_exc()[-1500:]}
    results[name + "_s"] = round(time.time() - t, 2)
    print(name, json.dumps(results[name])[:600], flush=True)
    with open(os.path.join(OUT, "check1.json"), "w") as f:
        json.dump(results, f, indent=1)


cases = [
    (128, 128, 64, 128, 0, 1), (256, 256, 64, 256, 0, 0), (128, 256, 256, 256, 0, 1), (1000, 512, 128, 256, 1, 0),
    (4096, 128, 512, 128, 2, 1), (4096, 256, 1024, 256, 2, 1), (64, 1024, 8192, 64, 1, 0), (64, 9, 256, 16, 0, 1),
    (16384, 2048, 512, 256, 1, 0), (300, 128, 64, 128, 0, 0),
]
for c in cases:
    run("gemm_%d_%d_%d_bn%d_e%d_f%d" % c, lambda c=c: gemm_case(*c))


def gemm_perf():
    M, N, K = 16384, 2048, 512
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    res = {}
    for epi in (0, 1):
        for _ in range(3):
            L.gdrn_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), None, None, _lib.ptr(out), M, N, K, epi, 0, 256,
                             _lib.current_stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        for _ in range(20):
            L.gdrn_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), None, None, _lib.ptr(out), M, N, K, epi, 0, 256,
                             _lib.current_stream())
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / 20
        res[f"epi{epi}_ms"] = ms
        res[f"epi{epi}_tflops"] = 2.0 * M * N * K / ms / 1e9
    # torch reference
    for _ in range(3):
        torch.matmul(A, W.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(20):
        torch.matmul(A, W.t())
    e1.record()
    sync()
    res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9
    return res


run("gemm_perf", gemm_perf)

# ---------------- model vs oracle ----------------
from oracle import gdrn_model_oracle as O  # noqa: E402  (scratch script: checker only)


def model_check():
    B = 4
    sd = make_state_dict()
    batch = make_batch(B=B, seed=3)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = O.gdrn_forward(sd, batch, return_maps=True, return_intermediate=True)
    model = GDRN_DoubleMask(default_cfg(with_maps=True))
    model.load_state_dict(sd)
    model.to(dev)
    gb = {k: v.to(dev) for k, v in batch.items()}
    out = model(gb["roi_img"], roi_classes=gb["roi_classes"], roi_coord_2d=gb["roi_coord_2d"], roi_cams=gb["roi_cams"],
                roi_centers=gb["roi_centers"], roi_whs=gb["roi_whs"], roi_extents=gb["roi_extents"],
                resize_ratios=gb["resize_ratios"], return_raw=True)
    sync()
    res = {}
    feat = model.debug_read("conv_feat", B, B * 64 * 1024).reshape(B, 8, 8, 1024).permute(0, 3, 1, 2).cpu()
    res["conv_feat_maxabs"] = float((feat - ref["conv_feat"]).abs().max())
    res["conv_feat_rel"] = float((feat - ref["conv_feat"]).norm() / ref["conv_feat"].norm())
    h64 = model.debug_read("head64", B, B * 4096 * 256)
    res["head64_absmean"] = float(h64.abs().mean())
    raw = out["raw"].cpu()
    res["rot6d_maxabs"] = float((raw[:, :6] - ref["rot6d"]).abs().max())
    res["t_maxabs"] = float((raw[:, 6:] - ref["t_"]).abs().max())
    res["rot6d_ref"] = ref["rot6d"][0].tolist()
    res["rot6d_got"] = raw[0, :6].tolist()
    for k in ("mask", "full_mask", "coor_x", "coor_y", "coor_z", "region"):
        res[k + "_maxabs"] = float((out[k].cpu() - ref[k]).abs().max())
    R, Rr = out["rot"].cpu().double(), ref["rot"].double()
    res["rot_err_rad_max"] = float((2 * torch.asin(((R - Rr).flatten(1).norm(dim=1) / (2 * 2 ** 0.5)).clamp(max=1.0))).max())
    res["trans_maxabs"] = float((out["trans"].cpu() - ref["trans"]).abs().max())
    return res


run("model_check", model_check)


def model_perf():
    B = 64
    sd = make_state_dict()
    batch = make_batch(B=B, seed=0)
    model = GDRN_DoubleMask(default_cfg())
    model.load_state_dict(sd)
    model.to(dev)
    gb = {k: v.to(dev) for k, v in batch.items()}
    kw = dict(roi_classes=gb["roi_classes"], roi_coord_2d=gb["roi_coord_2d"], roi_cams=gb["roi_cams"],
              roi_centers=gb["roi_centers"], roi_whs=gb["roi_whs"], roi_extents=gb["roi_extents"],
              resize_ratios=gb["resize_ratios"])
    for _ in range(3):
        model(gb["roi_img"], **kw)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = model(gb["roi_img"], **kw)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / 10
    return {"ms_per_batch64": ms, "rois_per_s": 64 / ms * 1e3, "finite": bool(torch.isfinite(out["rot"]).all())}


run("model_perf", model_perf)


# ---------------- op sanity ----------------
def ops_sanity():
    res = {}
    g = torch.Generator().manual_seed(0)
    # nnd vs cdist
    a = torch.rand(3, 500, 3, generator=g).to(dev)
    b = torch.rand(3, 700, 3, generator=g).to(dev)
    d1 = torch.empty(3, 500, device=dev); d2 = torch.empty(3, 700, device=dev)
    i1 = torch.empty(3, 500, dtype=torch.int32, device=dev); i2 = torch.empty(3, 700, dtype=torch.int32, device=dev)
    ok = L.nnd_forward_cuda(_lib.ptr(a), _lib.ptr(b), _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(i1), _lib.ptr(i2), 3, 500, 700,
                            _lib.current_stream())
    sync()
    cd = torch.cdist(a.double(), b.double()) ** 2
    res["nnd_ok"] = ok
    res["nnd_d1_err"] = float((d1.double() - cd.min(2)[0]).abs().max())
    res["nnd_i1_match"] = float((i1.long() == cd.argmin(2)).float().mean())
    res["nnd_d2_err"] = float((d2.double() - cd.min(1)[0]).abs().max())
    # fps: simple numpy check
    pts = (torch.rand(2, 3000, 3, generator=g) * 0.2 - 0.1)
    idx = torch.empty(2, 16, dtype=torch.int32, device=dev)
    rc = L.gdrn_fps_cuda(_lib.ptr(pts.to(dev)), _lib.ptr(idx), 3000, 16, 2, None, _lib.current_stream())
    sync()
    res["fps_rc"] = rc
    p = pts[0].numpy()
    c = (p.max(0) + p.min(0)) * np.float32(0.5)
    md = ((p - c) ** 2).sum(1)
    ref = []
    cur = int(md.argmax())
    for _ in range(16):
        ref.append(cur)
        d = ((p - p[cur]) ** 2).sum(1)
        md = np.minimum(md, d)
        md[ref] = -1
        cur = int(md.argmax())
    res["fps_got"] = idx[0].cpu().tolist()
    res["fps_ref"] = ref
    # raster sanity: sphere
    from gdrnpp_bop2022_b200.synthetic import make_icosphere_mesh
    v, f = make_icosphere_mesh(3, (0.1, 0.1, 0.1))
    vt, ft = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    pose = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0.5]], dtype=torch.float32, device=dev)[None]
    K = torch.tensor([[100, 0, 32], [0, 100, 32], [0, 0, 1]], dtype=torch.float32, device=dev)[None]
    depth = torch.empty(1, 64, 64, device=dev)
    scratch = torch.empty(64 * 64, dtype=torch.int64, device=dev)
    rc = L.rast_render_depth(_lib.ptr(vt), _lib.ptr(ft), v.shape[0], f.shape[0], _lib.ptr(pose), _lib.ptr(K), 1, 64, 64,
                             0.1, 100.0, 0, _lib.ptr(depth), None, _lib.ptr(scratch), _lib.current_stream())
    sync()
    res["rast_rc"] = rc
    res["rast_center_depth"] = float(depth[0, 32, 32])
    res["rast_cov"] = int((depth > 0).sum())
    return res


run("ops_sanity", ops_sanity)
print("DONE")

"""ORACLE (test infrastructure, NOT product code): ctypes front-end of liboracle_ops.so (ops_oracle.c) plus
numpy restatements of the rasteriser / depth refine / uncertainty-PnP reference code.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle_ops.so")
        if not os.path.exists(path):
            from . import build_ref

            build_ref.build_oracle_ops()
        _lib = ctypes.CDLL(path)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fps(pts, sn, start=-1):
    pts = np.ascontiguousarray(pts, np.float32)
    idx = np.zeros(sn, np.int32)
    lib().oracle_fps(_p(pts), _p(idx), pts.shape[0], sn, int(start))
    return idx


def generate_hypothesis(direct, coords, idxs, vanishing_point=False):
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    idxs = np.ascontiguousarray(idxs, np.int32)
    tn, vn, _ = direct.shape
    hn = idxs.shape[0]
    hypo = np.zeros((hn, vn, 3 if vanishing_point else 2), np.float32)
    fn = lib().oracle_generate_hypothesis_vp if vanishing_point else lib().oracle_generate_hypothesis
    fn(_p(direct), _p(coords), _p(idxs), _p(hypo), tn, vn, hn)
    return hypo


def voting(direct, coords, hypo, thresh, vanishing_point=False, inliers=None):
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    hypo = np.ascontiguousarray(hypo, np.float32)
    tn, vn, _ = direct.shape
    hn = hypo.shape[0]
    if inliers is None:
        inliers = np.zeros((hn, vn, tn), np.uint8)
    counts = np.zeros((hn, vn), np.int32)
    lib().oracle_voting(_p(direct), _p(coords), _p(hypo), _p(inliers), _p(counts), tn, vn, hn,
                        ctypes.c_float(thresh), int(vanishing_point))
    return inliers, counts


def nnd_forward(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    bs, n, _ = a.shape
    m = b.shape[1]
    d1 = np.zeros((bs, n), np.float32); i1 = np.zeros((bs, n), np.int32)
    d2 = np.zeros((bs, m), np.float32); i2 = np.zeros((bs, m), np.int32)
    lib().oracle_nnd_forward(_p(a), _p(b), _p(d1), _p(i1), bs, n, m)
    lib().oracle_nnd_forward(_p(b), _p(a), _p(d2), _p(i2), bs, m, n)
    return d1, d2, i1, i2


def nnd_backward(a, b, g1, g2, i1, i2):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    bs, n, _ = a.shape
    m = b.shape[1]
    ga = np.zeros_like(a); gb = np.zeros_like(b)
    lib().oracle_nnd_backward(_p(a), _p(b), _p(np.ascontiguousarray(g1, np.float32)), _p(np.ascontiguousarray(i1, np.int32)), _p(ga), _p(gb), bs, n, m)
    lib().oracle_nnd_backward(_p(b), _p(a), _p(np.ascontiguousarray(g2, np.float32)), _p(np.ascontiguousarray(i2, np.int32)), _p(gb), _p(ga), bs, m, n)
    return ga, gb


def flow(depth_src, depth_tgt, KT, Kinv):
    ds = np.ascontiguousarray(depth_src, np.float32); dt = np.ascontiguousarray(depth_tgt, np.float32)
    KT = np.ascontiguousarray(KT, np.float32); Kinv = np.ascontiguousarray(Kinv, np.float32)
    B, _, H, W = ds.shape
    fl = np.zeros((B, 2, H, W), np.float32); va = np.zeros((B, 1, H, W), np.float32)
    lib().oracle_flow(_p(ds), _p(dt), _p(KT), _p(Kinv), _p(fl), _p(va), B, H, W)
    return fl, va


# ---------------------------------------------------------------------------------------------------------
# Rasteriser: numpy z-buffer following the GL conventions of lib/render_vispy/renderer.py:461-476 (projection),
# :155-182 (read-back, flip, depth decode) -- see SURVEY.md Appendix B.  Pixel (r,c) sampled at (c+.5, r+.5).
def render_depth(verts, faces, pose, K, H, W, znear=0.1, zfar=100.0):
    v = verts.astype(np.float64) @ pose[:3, :3].astype(np.float64).T + pose[:3, 3].astype(np.float64)
    K = K.astype(np.float64)
    z = v[:, 2]
    u = (K[0, 0] * v[:, 0] + K[0, 1] * v[:, 1]) / z + K[0, 2]
    w = K[1, 1] * v[:, 1] / z + K[1, 2]
    depth = np.full((H, W), np.inf)
    for f in faces:
        zz = z[f]
        if (zz < znear).any():
            continue
        uu, vv = u[f], w[f]
        area = (uu[1] - uu[0]) * (vv[2] - vv[0]) - (uu[2] - uu[0]) * (vv[1] - vv[0])
        if area == 0:
            continue
        c0 = max(0, int(math.floor(uu.min() - 0.5))); c1 = min(W - 1, int(math.ceil(uu.max() - 0.5)))
        r0 = max(0, int(math.floor(vv.min() - 0.5))); r1 = min(H - 1, int(math.ceil(vv.max() - 0.5)))
        if c1 < c0 or r1 < r0:
            continue
        xs = np.arange(c0, c1 + 1) + 0.5
        ys = np.arange(r0, r1 + 1) + 0.5
        sx, sy = np.meshgrid(xs, ys)
        w0 = ((uu[1] - sx) * (vv[2] - sy) - (uu[2] - sx) * (vv[1] - sy)) / area
        w1 = ((uu[2] - sx) * (vv[0] - sy) - (uu[0] - sx) * (vv[2] - sy)) / area
        w2 = 1.0 - w0 - w1
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
        if not inside.any():
            continue
        zp = 1.0 / (w0 / zz[0] + w1 / zz[1] + w2 / zz[2])
        ok = inside & (zp >= znear) & (zp <= zfar)
        sub = depth[r0:r1 + 1, c0:c1 + 1]
        sub[ok] = np.minimum(sub[ok], zp[ok])
    depth[~np.isfinite(depth)] = 0.0
    return depth.astype(np.float32)


# Depth refine: core/gdrn_modeling/engine/gdrn_evaluator.py:521-561 (one iteration, render supplied).
def depth_refine_step(xyz_hw3, mask_hw, depth_sensor, ren_dp, K_crop, trans, thresh=0.8):
    import torch

    crop_res = mask_hw.shape[0]
    ren_mask = ren_dp > 0
    sensor_mask = depth_sensor > 0
    q = torch.norm(torch.from_numpy(xyz_hw3), dim=-1) * torch.from_numpy(mask_hw)
    q = q.numpy() * ren_mask * sensor_mask
    norm_sum = q.sum()
    if norm_sum == 0:
        return trans.copy()
    q = q / norm_sum
    norm_mask = q > (q.max() * thresh)
    yy, xx = np.argwhere(norm_mask).T
    depth_diff = depth_sensor[yy, xx] - ren_dp[yy, xx]
    depth_adjustment = np.median(depth_diff)
    yx_coords = np.meshgrid(np.arange(crop_res), np.arange(crop_res))
    yx_coords = np.stack(yx_coords[::-1], axis=-1)
    yx_ray_2d = (yx_coords * q[..., None]).sum(axis=(0, 1))
    ray_3d = np.linalg.inv(K_crop) @ (*yx_ray_2d[::-1], 1)
    ray_3d /= ray_3d[2]
    return trans + (ray_3d[:, None] * depth_adjustment).reshape(3)


# Uncertainty PnP: core/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:16-34 residual; Ceres 2.0 default
# trust-region LM restated (libceres absent -> PARITY UNPINNED against Ceres proper; known-answer recipe
# of the reference's own main() :98-156 is the anchor).
def _rodrigues_point(aa, p):
    th2 = float(aa @ aa)
    if th2 > np.finfo(np.float64).eps:
        th = math.sqrt(th2)
        w = aa / th
        c, s = math.cos(th), math.sin(th)
        return p * c + np.cross(w, p) * s + w * (w @ p) * (1 - c)
    return p + np.cross(aa, p)


def upnp_residuals(x, p2, p3, w, K):
    res = np.zeros(2 * len(p2))
    for i in range(len(p2)):
        q = _rodrigues_point(x[:3], p3[i]) + x[3:]
        dx = K[0, 0] * q[0] / q[2] + K[0, 2] - p2[i, 0]
        dy = K[1, 1] * q[1] / q[2] + K[1, 2] - p2[i, 1]
        res[2 * i] = w[i, 0] * dx + w[i, 1] * dy
        res[2 * i + 1] = w[i, 1] * dx + w[i, 2] * dy
    return res


def uncertainty_pnp(p2, p3, w, K, init_rt, max_iter=50):
    x = np.asarray(init_rt, np.float64).copy()

    def jac(x):
        J = np.zeros((2 * len(p2), 6))
        for k in range(6):  # central differences in fp64 (checker: tolerance-level agreement only)
            h = 1e-7 * max(1.0, abs(x[k]))
            xp, xm = x.copy(), x.copy()
            xp[k] += h; xm[k] -= h
            J[:, k] = (upnp_residuals(xp, p2, p3, w, K) - upnp_residuals(xm, p2, p3, w, K)) / (2 * h)
        return J

    r = upnp_residuals(x, p2, p3, w, K)
    cost = 0.5 * r @ r
    J = jac(x)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    radius, dec = 1e4, 2.0
    for _ in range(max_iter):
        Js = J * scale
        g = Js.T @ r
        A = Js.T @ Js
        if np.abs(J.T @ r).max() <= 1e-10:
            break
        D = np.clip(np.diag(A), 1e-6, 1e32) / radius
        try:
            step = np.linalg.solve(A + np.diag(D), -g)
        except np.linalg.LinAlgError:
            radius /= dec; dec *= 2
            continue
        mc = -step @ (g + 0.5 * A @ step)
        delta = step * scale
        ok = False
        if mc > 0:
            if np.linalg.norm(delta) <= 1e-8 * (np.linalg.norm(x) + 1e-8):
                break
            xn = x + delta
            rn = upnp_residuals(xn, p2, p3, w, K)
            cn = 0.5 * rn @ rn
            rel = (cost - cn) / mc
            if rel > 1e-3:
                ok = True
                change = cost - cn
                prev = cost
                x, r, cost = xn, rn, cn
                J = jac(x)
                radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2 * rel - 1) ** 3))
                dec = 2.0
                if abs(change) <= 1e-6 * prev:
                    break
        if not ok:
            radius /= dec
            dec *= 2
            if radius < 1e-32:
                break
    return x


# ---- ROI crop + resize: restatement of cv2.warpAffine (ops_oracle.c), pinned against cv2 in tests/test_oracle_pinning.py
def warp_affine_u8(img, M, out_size):
    """img [H,W,C] uint8, M [2,3] float64 forward transform, out_size (w, h) -> [h,w,C] uint8 (INTER_LINEAR, border 0)."""
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    ow, oh = out_size
    dst = np.zeros((oh, ow, img.shape[2]), np.uint8)
    lib().oracle_warp_affine_u8(_p(img), img.shape[0], img.shape[1], img.shape[2], _p(M), _p(dst), oh, ow)
    return dst


def warp_affine_f32(src, M, out_size, nearest=False):
    src = np.ascontiguousarray(src, np.float32)
    if src.ndim == 2:
        src = src[:, :, None]
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    ow, oh = out_size
    dst = np.zeros((oh, ow, src.shape[2]), np.float32)
    lib().oracle_warp_affine_f32(_p(src), src.shape[0], src.shape[1], src.shape[2], _p(M), _p(dst), oh, ow, int(bool(nearest)))
    return dst


def crop_resize_roi(image_u8, M, out_res, pixel_mean=(0.0, 0.0, 0.0), pixel_std=(255.0, 255.0, 255.0)):
    """crop_resize_by_warp_affine + transpose + normalize_image + astype(float32) (predictor_gdrn.py:417-422)."""
    crop = warp_affine_u8(image_u8, M, (out_res, out_res)).transpose(2, 0, 1)
    mean = np.array(pixel_mean, np.float64).reshape(-1, 1, 1)
    std = np.array(pixel_std, np.float64).reshape(-1, 1, 1)
    return ((crop - mean) / std).astype(np.float32)

"""Batched GPU RANSAC-PnP behind the reference's TEST.USE_PNP surfaces.

Mirrors ``get_pnp_ransac_pose`` / ``get_img_model_points_with_coords2d`` (core/gdrn_modeling/engine/gdrn_evaluator.py:
1122-1221) and ``misc.pnp_v2(..., ransac=True, ransac_reprojErr=3, ransac_iter=100)`` (lib/pysixd/misc.py:153-208), for all
ROIs of a batch in one launch of libgdrn_b200.so (csrc/pnp_ransac.cu).  No CPU fallback.
"""
import torch

from . import _lib


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


@_lib.on_device(0)
def pnp_ransac_from_maps(coor_x, coor_y, coor_z, mask, roi_coord_2d, im_H, im_W, roi_extents, cams, mask_thr=0.5,
                         reproj_err=3.0, iters=100, seed=0, idxs=None, return_inliers=False):
    """coor_x/y/z, mask: [n,1,hw,hw] raw outputs of GDRN_DoubleMask.forward; roi_coord_2d [n,2,hw,hw]; im_H, im_W [n];
    roi_extents [n,3]; cams [n,3,3] -> poses [n,3,4] (R|t; -100 where the reference returns its sentinel)."""
    if not coor_x.is_cuda:
        raise _lib.GdrnError("pnp_ransac_from_maps needs CUDA tensors (no CPU fallback)")
    dev = coor_x.device
    n, hw = coor_x.shape[0], coor_x.shape[-1]
    cx, cy, cz, m, c2d = (_f32(t, dev) for t in (coor_x, coor_y, coor_z, mask, roi_coord_2d))
    im_hw = torch.stack([torch.as_tensor(im_H, dtype=torch.float32).reshape(-1), torch.as_tensor(im_W, dtype=torch.float32).reshape(-1)], dim=1)
    im_hw = _f32(im_hw, dev)
    ext, Ks = _f32(roi_extents, dev), _f32(cams, dev).reshape(-1, 3, 3)
    if Ks.shape[0] == 1 and n > 1:
        Ks = Ks.expand(n, 3, 3).contiguous()
    poses = torch.empty((n, 3, 4), dtype=torch.float32, device=dev)
    ninl = torch.empty((n,), dtype=torch.int32, device=dev)
    imask = torch.empty((n, hw * hw), dtype=torch.uint8, device=dev) if return_inliers else None
    ix = None if idxs is None else idxs.detach().to(device=dev, dtype=torch.int32).contiguous()
    if ix is not None:
        iters = ix.shape[1]
    if n > 0:
        _lib.check(_lib.lib().gdrn_pnp_ransac_maps(_lib.ptr(cx), _lib.ptr(cy), _lib.ptr(cz), _lib.ptr(m), _lib.ptr(c2d), _lib.ptr(im_hw),
                                                   _lib.ptr(ext), _lib.ptr(Ks), _lib.ptr(ix), n, hw, int(iters), float(mask_thr),
                                                   float(reproj_err), int(seed) & 0xFFFFFFFF, _lib.ptr(poses), _lib.ptr(ninl),
                                                   _lib.ptr(imask), _lib.current_stream()), "gdrn_pnp_ransac_maps")
    return (poses, ninl, imask.view(n, hw, hw)) if return_inliers else poses


@_lib.on_device(0)
def solve_pnp_ransac(pts3d, pts2d, cams, reproj_err=3.0, iters=100, seed=0, idxs=None, return_inliers=False):
    """Explicit correspondences: pts3d [n,npts,3], pts2d [n,npts,2], cams [n,3,3] (CUDA) -> poses [n,3,4]
    (the batched stand-in for n calls of misc.pnp_v2(points_3d, points_2d, K, ransac=True))."""
    if not pts3d.is_cuda:
        raise _lib.GdrnError("solve_pnp_ransac needs CUDA tensors (no CPU fallback)")
    dev = pts3d.device
    n, npts = pts3d.shape[0], pts3d.shape[1]
    p3, p2, Ks = _f32(pts3d, dev), _f32(pts2d, dev), _f32(cams, dev).reshape(-1, 3, 3)
    if Ks.shape[0] == 1 and n > 1:
        Ks = Ks.expand(n, 3, 3).contiguous()
    poses = torch.empty((n, 3, 4), dtype=torch.float32, device=dev)
    ninl = torch.empty((n,), dtype=torch.int32, device=dev)
    imask = torch.empty((n, npts), dtype=torch.uint8, device=dev) if return_inliers else None
    ix = None if idxs is None else idxs.detach().to(device=dev, dtype=torch.int32).contiguous()
    if ix is not None:
        iters = ix.shape[1]
    _lib.check(_lib.lib().gdrn_pnp_ransac_points(_lib.ptr(p3), _lib.ptr(p2), _lib.ptr(Ks), _lib.ptr(ix), n, npts, int(iters),
                                                 float(reproj_err), int(seed) & 0xFFFFFFFF, _lib.ptr(poses), _lib.ptr(ninl),
                                                 _lib.ptr(imask), _lib.current_stream()), "gdrn_pnp_ransac_points")
    return (poses, ninl, imask) if return_inliers else poses

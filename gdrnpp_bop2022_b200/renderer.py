"""CUDA depth rasteriser + fast depth refinement behind the reference's renderer / evaluator surfaces.

Mirrors:
  * lib/render_vispy/renderer.py  Renderer(size, cam).{clear, set_cam, draw_model, finish} -> (rgb, depth)
    (the renderer fast depth refine uses: engine/gdrn_evaluator.py:64-84,520-526; demo/predictor_gdrn.py:100-109)
  * lib/egl_renderer/egl_renderer_v3.py  EGLRenderer.render(obj_ids, poses, K=, pc_cam_tensor=) (depth = pc_cam[...,2])
  * GDRN_Evaluator.process_depth_refine (engine/gdrn_evaluator.py:461-573) == GdrnPredictor.process_depth_refine
    (demo/predictor_gdrn.py:195-286), batched over ROIs on the GPU.
  * get_out_mask / get_out_coor (engine/engine_utils.py:295-333), get_K_crop_resize (core/utils/camera_geometry.py:6-21)
No GL / EGL anywhere; no CPU fallback.
"""
import numpy as np
import torch

from . import _lib


def _f32c(t, dev=None):
    return t.detach().to(device=dev or t.device, dtype=torch.float32).contiguous()


@_lib.on_device(0)
def render_depth(verts, faces, poses, Ks, H, W, znear=0.1, zfar=100.0, quantize_bits=0, return_xyz=False):
    """verts [V,3] f32, faces [F,3] i32, poses [n,3,4], Ks [n,3,3] (CUDA) -> depth [n,H,W] (0 = background)."""
    if not verts.is_cuda:
        raise _lib.GdrnError("render_depth needs CUDA tensors (no CPU fallback)")
    dev = verts.device
    verts = _f32c(verts)
    faces = faces.detach().to(device=dev, dtype=torch.int32).contiguous()
    poses = _f32c(poses, dev).reshape(-1, 3, 4)
    Ks = _f32c(Ks, dev).reshape(-1, 3, 3)
    n = poses.shape[0]
    if Ks.shape[0] == 1 and n > 1:
        Ks = Ks.expand(n, 3, 3).contiguous()
    depth = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    xyz = torch.empty((n, H, W, 3), dtype=torch.float32, device=dev) if return_xyz else None
    scratch = torch.empty((n * H * W,), dtype=torch.int64, device=dev)
    L = _lib.lib()
    _lib.check(L.rast_render_depth(_lib.ptr(verts), _lib.ptr(faces), verts.shape[0], faces.shape[0], _lib.ptr(poses),
                                   _lib.ptr(Ks), n, H, W, float(znear), float(zfar), int(quantize_bits), _lib.ptr(depth),
                                   _lib.ptr(xyz), _lib.ptr(scratch), _lib.current_stream()), "rast_render_depth")
    return (depth, xyz) if return_xyz else depth


def get_out_mask(pred_mask, mask_loss_type="L1"):
    """engine_utils.py:313-333 (L1: per-ROI min-max normalisation to [0,1])."""
    bs = pred_mask.shape[0]
    if mask_loss_type == "L1":
        mx = torch.max(pred_mask.view(bs, -1), dim=-1)[0].view(bs, 1, 1, 1)
        mn = torch.min(pred_mask.view(bs, -1), dim=-1)[0].view(bs, 1, 1, 1)
        return (pred_mask - mn) / (mx - mn)
    if mask_loss_type in ("BCE", "RW_BCE", "dice"):
        return torch.sigmoid(pred_mask)
    raise NotImplementedError(mask_loss_type)


def get_out_coor(coor_x, coor_y, coor_z):
    """engine_utils.py:295-310 (regression branch)."""
    return torch.cat([coor_x, coor_y, coor_z], dim=1)


def get_K_crop_resize(K, crop_xy, resize_ratio):
    """core/utils/camera_geometry.py:6-21."""
    bs = K.shape[0]
    new_K = K.clone()
    new_K[:, [0, 1], 2] = K[:, [0, 1], 2] - crop_xy
    new_K[:, [0, 1]] = new_K[:, [0, 1]] * resize_ratio.view(bs, -1, 1)
    return new_K


@_lib.on_device(2)
def depth_refine(verts, faces, rot, trans, K_crop, xyz, mask, depth_sensor, iters=2, thresh=0.8, mesh_ids=None,
                 znear=0.1, zfar=100.0, mask_loss_type="L1"):
    """Batched fast depth refine (gdrn_evaluator.py:515-561): `iters` x (render depth at the current pose ->
    weighted-median depth offset along the weighted-centroid ray).

    verts/faces: one mesh (tensors) or lists of meshes with `mesh_ids` [n] selecting per ROI.
    rot [n,3,3], trans [n,3], K_crop [n,3,3], xyz [n,3,h,w] (coor maps), mask [n,1,h,w] raw visible-mask output,
    depth_sensor [n,h,w] metres.  Returns refined trans [n,3]."""
    dev = rot.device
    n = rot.shape[0]
    hw = xyz.shape[-1]
    rot = _f32c(rot)
    t = _f32c(trans).clone()
    K_crop = _f32c(K_crop)
    xyz = _f32c(xyz)
    mnorm = _f32c(get_out_mask(mask, mask_loss_type)).reshape(n, hw, hw)
    sensor = _f32c(depth_sensor).reshape(n, hw, hw)
    L = _lib.lib()
    meshes = [(verts, faces)] if not isinstance(verts, (list, tuple)) else list(zip(verts, faces))
    ids = torch.zeros(n, dtype=torch.long) if mesh_ids is None else torch.as_tensor(mesh_ids).cpu().long()
    for _ in range(iters):
        poses = torch.cat([rot, t[:, :, None]], dim=2).contiguous()
        ren = torch.zeros((n, hw, hw), dtype=torch.float32, device=dev)
        for mi, (v, f) in enumerate(meshes):
            sel = torch.nonzero(ids == mi).flatten().to(dev)
            if sel.numel() == 0:
                continue
            ren[sel] = render_depth(v, f, poses[sel], K_crop[sel], hw, hw, znear, zfar)
        _lib.check(L.gdrn_depth_refine_step(_lib.ptr(xyz), _lib.ptr(mnorm), _lib.ptr(sensor), _lib.ptr(ren),
                                            _lib.ptr(K_crop), _lib.ptr(t), n, hw, float(thresh), _lib.current_stream()),
                   "gdrn_depth_refine_step")
    return t


class Model3D:
    """Minimal stand-in for lib/render_vispy/model3d.py objects: vertices in metres + triangle indices."""

    def __init__(self, vertices, faces, device="cuda"):
        self.vertices = torch.as_tensor(np.asarray(vertices, np.float32)).to(device)
        self.faces = torch.as_tensor(np.asarray(faces, np.int32)).to(device)


class Renderer:
    """lib/render_vispy/renderer.py:Renderer surface (depth only; rgb is returned as zeros)."""

    def __init__(self, size, cam, model_paths=None, scale_to_meter=1.0, gpu_id=None, device="cuda"):
        self.width, self.height = size
        self.shape = (self.height, self.width)
        self.device = torch.device(device)
        self.set_cam(cam)
        self._draws = []

    def set_cam(self, cam, clip_near=0.1, clip_far=100.0):
        self.cam = np.asarray(cam, np.float32)
        self.clip_near, self.clip_far = clip_near, clip_far

    def clear(self, color=True, depth=True):
        self._draws = []

    def draw_model(self, model, pose, ambient_weight=0.5, light=(0, 0, 1), light_col=(1, 1, 1)):
        self._draws.append((model, np.asarray(pose, np.float32)[:3, :4]))

    def finish(self, only_color=False, to_255=False):
        rgb = np.zeros(self.shape + (3,), np.uint8 if to_255 else np.float32)
        if only_color:
            return rgb
        depth = None
        K = torch.from_numpy(self.cam)[None].to(self.device)
        for model, pose in self._draws:
            d = render_depth(model.vertices, model.faces, torch.from_numpy(pose)[None].to(self.device), K, self.height,
                             self.width, self.clip_near, self.clip_far)[0]
            if depth is None:
                depth = d
            else:  # nearest surface wins across draws (shared z-buffer in GL)
                depth = torch.where((d > 0) & ((depth == 0) | (d < depth)), d, depth)
        if depth is None:
            depth = torch.zeros(self.shape, device=self.device)
        return rgb, depth.cpu().numpy()


class EGLRenderer:
    """lib/egl_renderer/egl_renderer_v3.py:EGLRenderer surface for the camera-space point-cloud / depth outputs."""

    def __init__(self, models, K=None, width=640, height=480, znear=0.25, zfar=6.0, device="cuda", **_):
        self.models = list(models)  # Model3D-like objects (the reference takes model paths; PLY loading is out of scope)
        self.K = None if K is None else np.asarray(K, np.float32)
        self.width, self.height, self.znear, self.zfar = width, height, znear, zfar
        self.device = torch.device(device)

    def render(self, obj_ids, poses, K=None, pc_cam_tensor=None, seg_tensor=None, **_):
        """Writes camera-space xyz into pc_cam_tensor[..., :3] ([H,W,4] float CUDA, like the reference) and returns depth."""
        Kuse = self.K if K is None else np.asarray(K, np.float32)
        Kt = torch.from_numpy(Kuse)[None].to(self.device)
        depth = torch.zeros((self.height, self.width), device=self.device)
        xyz = torch.zeros((self.height, self.width, 3), device=self.device)
        seg = torch.zeros((self.height, self.width), device=self.device)
        for k, (oid, pose) in enumerate(zip(obj_ids, poses)):
            m = self.models[oid]
            p = torch.as_tensor(np.asarray(pose, np.float32)[:3, :4])[None].to(self.device)
            d, x = render_depth(m.vertices, m.faces, p, Kt, self.height, self.width, self.znear, self.zfar, return_xyz=True)
            upd = (d[0] > 0) & ((depth == 0) | (d[0] < depth))
            depth = torch.where(upd, d[0], depth)
            xyz = torch.where(upd[..., None], x[0], xyz)
            seg = torch.where(upd, torch.full_like(seg, float(k + 1)), seg)
        if pc_cam_tensor is not None:
            pc_cam_tensor[..., :3] = xyz
            if pc_cam_tensor.shape[-1] > 3:
                pc_cam_tensor[..., 3] = (depth > 0).float()
        if seg_tensor is not None:
            seg_tensor[..., 0] = seg
        return depth

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


_MESH_IDS = {}   # (verts.data_ptr, faces.data_ptr, V, F) -> id in the library's mesh registry (rast_upload_mesh)


def upload_mesh(verts, faces):
    """Register a mesh with the library once (rast_upload_mesh keeps its own device copy) and return its id.  Tensors
    are identified by (data_ptr, shape): re-uploading the same tensors is free; a Model3D caches its id itself."""
    key = (verts.data_ptr(), faces.data_ptr(), int(verts.shape[0]), int(faces.shape[0]))
    mid = _MESH_IDS.get(key)
    if mid is None:
        v = _f32c(verts)
        f = faces.detach().to(dtype=torch.int32).contiguous()
        mid = _lib.lib().rast_upload_mesh(_lib.ptr(v), int(v.shape[0]), _lib.ptr(f), int(f.shape[0]))
        if mid < 0:
            raise _lib.GdrnError(f"rast_upload_mesh failed (code {mid}): {_lib.last_error()}")
        _MESH_IDS[key] = mid
    return mid


@_lib.on_device(1)
def render_meshes(mesh_ids, poses, Ks, H, W, znear=0.1, zfar=100.0, quantize_bits=0, return_xyz=False):
    """One render per ROI of the REGISTERED mesh `mesh_ids[i]` (int32 CUDA tensor of registry ids): poses [n,3,4],
    Ks [n,3,3] -> depth [n,H,W].  The multi-object form of the reference's per-object draw loop, one launch, no sync."""
    dev = poses.device
    poses = _f32c(poses).reshape(-1, 3, 4)
    n = poses.shape[0]
    Ks = _f32c(Ks, dev).reshape(-1, 3, 3)
    if Ks.shape[0] == 1 and n > 1:
        Ks = Ks.expand(n, 3, 3).contiguous()
    ids = mesh_ids.detach().to(device=dev, dtype=torch.int32).contiguous()
    depth = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    xyz = torch.empty((n, H, W, 3), dtype=torch.float32, device=dev) if return_xyz else None
    scratch = torch.empty((n * H * W,), dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().rast_render_meshes(_lib.ptr(ids), _lib.ptr(poses), _lib.ptr(Ks), n, H, W, float(znear), float(zfar),
                                             int(quantize_bits), _lib.ptr(depth), _lib.ptr(xyz), _lib.ptr(scratch),
                                             _lib.current_stream()), "rast_render_meshes")
    return (depth, xyz) if return_xyz else depth


@_lib.on_device(2)
def depth_refine(verts, faces, rot, trans, K_crop, xyz, mask, depth_sensor, iters=2, thresh=0.8, mesh_ids=None,
                 znear=0.1, zfar=100.0, mask_loss_type="L1"):
    """Batched fast depth refine (gdrn_evaluator.py:515-561): `iters` x (render depth at the current pose ->
    weighted-median depth offset along the weighted-centroid ray).

    verts/faces: one mesh (tensors) or lists of meshes with `mesh_ids` [n] selecting per ROI.
    rot [n,3,3], trans [n,3], K_crop [n,3,3], xyz [n,3,h,w] (coor maps), mask [n,1,h,w] RAW visible-mask output
    (get_out_mask -- per-ROI min-max for the L1 head, sigmoid otherwise -- runs inside the refine kernel),
    depth_sensor [n,h,w] metres.  Returns refined trans [n,3].  Per iteration: one render launch over all ROIs (every
    ROI picks its mesh from the library's registry by id) + one refine launch; no host loop over meshes, no sync."""
    dev = rot.device
    n = rot.shape[0]
    hw = xyz.shape[-1]
    rot = _f32c(rot)
    t = _f32c(trans).clone()
    K_crop = _f32c(K_crop)
    xyz = _f32c(xyz)
    raw_mask = _f32c(mask).reshape(n, hw, hw)
    if mask_loss_type == "L1":
        mask_mode = 1
    elif mask_loss_type in ("BCE", "RW_BCE", "dice"):
        mask_mode = 2
    else:
        raise NotImplementedError(mask_loss_type)
    sensor = _f32c(depth_sensor).reshape(n, hw, hw)
    L = _lib.lib()
    meshes = [(verts, faces)] if not isinstance(verts, (list, tuple)) else list(zip(verts, faces))
    reg = torch.tensor([upload_mesh(v, f) for v, f in meshes], dtype=torch.int32, device=dev)   # list index -> registry id
    if mesh_ids is None:
        ids = reg[:1].expand(n).contiguous()
    else:
        ids = reg[torch.as_tensor(mesh_ids).to(device=dev, dtype=torch.long)].contiguous()
    for _ in range(iters):
        poses = torch.cat([rot, t[:, :, None]], dim=2).contiguous()
        ren = render_meshes(ids, poses, K_crop, hw, hw, znear, zfar)
        _lib.check(L.gdrn_depth_refine_step_ex(_lib.ptr(xyz), _lib.ptr(raw_mask), mask_mode, _lib.ptr(sensor), _lib.ptr(ren),
                                               _lib.ptr(K_crop), _lib.ptr(t), n, hw, float(thresh), _lib.current_stream()),
                   "gdrn_depth_refine_step_ex")
    return t


class Model3D:
    """Minimal stand-in for lib/render_vispy/model3d.py objects: vertices in metres + triangle indices."""

    def __init__(self, vertices, faces, device="cuda"):
        self.vertices = torch.as_tensor(np.asarray(vertices, np.float32)).to(device)
        self.faces = torch.as_tensor(np.asarray(faces, np.int32)).to(device)


def load_models(model_paths, scale_to_meter=1.0, device="cuda", **_):
    """lib/render_vispy/model3d.py:load_models surface: PLY paths -> Model3D list (vertices * scale_to_meter)."""
    from .ply import load_ply

    out = []
    for p in model_paths:
        m = load_ply(p, vertex_scale=scale_to_meter)
        if "faces" not in m:
            raise ValueError(f"{p}: the rasteriser needs a triangle mesh (no face element)")
        out.append(Model3D(m["pts"], m["faces"], device=device))
    return out


class Renderer:
    """lib/render_vispy/renderer.py:Renderer surface (:78-182): Renderer(size, cam, model_paths=, scale_to_meter=),
    set_cam / clear / draw_model / finish -> (rgb, depth).  Depth comes from the CUDA rasteriser with the vispy path's
    z-buffer read-back emulated (fixed-point window depth, decode mult / (d + addi), :176-182); rgb is not rendered
    (depth refine, the only caller on the hot path, ignores it) and is returned as zeros."""

    def __init__(self, size, cam, model_paths=None, scale_to_meter=1.0, gpu_id=None, device="cuda", depth_bits=24):
        self.width, self.height = size
        self.size = tuple(size)
        self.shape = (self.height, self.width)
        self.device = torch.device(device if gpu_id is None else "cuda:%d" % gpu_id)
        self.depth_bits = depth_bits      # GL depth attachment precision emulated by finish() (0 = float z)
        self.models = None
        if model_paths is not None:
            self.models = load_models(model_paths, scale_to_meter=scale_to_meter, device=self.device)
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
                             self.width, self.clip_near, self.clip_far, quantize_bits=self.depth_bits)[0]
            if depth is None:
                depth = d
            else:  # nearest surface wins across draws (shared z-buffer in GL)
                depth = torch.where((d > 0) & ((depth == 0) | (d < depth)), d, depth)
        if depth is None:
            depth = torch.zeros(self.shape, device=self.device)
        return rgb, depth.cpu().numpy()


class EGLRenderer:
    """lib/egl_renderer/egl_renderer_v3.py:EGLRenderer surface (:838-857, 1183-1228) for the geometry outputs:
    EGLRenderer(model_paths, K=, width=, height=, vertex_scale=, znear=, zfar=) and
    render(obj_ids, poses, K=, pc_cam_tensor=, seg_tensor=) writing camera-space xyz (float-exact per fragment, like the
    interpolated varying of shader_textureless_texture.vs:36) into the caller's [H,W,4] CUDA tensor.  `model_paths` may
    also be Model3D-like objects.  Colour / normal / object-space outputs are not rendered."""

    def __init__(self, model_paths, K=None, width=640, height=480, vertex_scale=1.0, znear=0.25, zfar=6.0, device="cuda", **_):
        ms = list(model_paths)
        if ms and isinstance(ms[0], (str, bytes)) or (ms and hasattr(ms[0], "__fspath__")):
            self.models = load_models(ms, scale_to_meter=vertex_scale, device=device)
        else:
            self.models = ms
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

"""Host-side mirror of the reference model surface for the hot path.

Reference interface mirrored (paths relative to the reference tree):
  * ``build_model_optimizer(cfg, is_test)`` -> ``(model, optimizer|None)``
    (core/gdrn_modeling/models/GDRN_double_mask.py:539-615, called from main_gdrn.py:158)
  * ``GDRN_DoubleMask.forward(x, roi_classes=, roi_coord_2d=, roi_cams=, roi_centers=, roi_whs=,
    roi_extents=, resize_ratios=, do_loss=False)`` -> ``{"rot", "trans"[, "mask", "full_mask",
    "coor_x", "coor_y", "coor_z", "region"]}`` (GDRN_double_mask.py:66-214)
  * ``state_dict`` key names / shapes of a reference checkpoint, so
    ``MyCheckpointer(model).resume_or_load`` (core/utils/my_checkpoint.py:35-83) works unchanged.

All arithmetic happens in libgdrn_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/gdrn_b200.h).  There is no PyTorch/CPU fallback: without the library or a CUDA device the
forward raises.
"""
import ctypes
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib
from .synthetic import CONVNEXT_ARCH, make_state_dict


def _cfg_get(cfg, path, default=None):
    cur = cfg
    for part in path.split("."):
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(part, None)
        else:
            cur = getattr(cur, part, None)
    return default if cur is None else cur


def default_cfg(num_classes=21, arch="convnext_base", with_maps=False):
    """Minimal stand-in for the mmcv Config of convnext_a6_AugCosyAAEGray_..._ycbv.py."""
    return SimpleNamespace(
        MODEL=SimpleNamespace(
            POSE_NET=SimpleNamespace(
                NAME="GDRN_double_mask",
                NUM_CLASSES=num_classes,
                OUTPUT_RES=64,
                BACKBONE=SimpleNamespace(INIT_CFG=SimpleNamespace(type="timm/" + arch)),
                GEO_HEAD=SimpleNamespace(NUM_REGIONS=64, XYZ_CLASS_AWARE=True, MASK_CLASS_AWARE=True,
                                         REGION_CLASS_AWARE=True),
                PNP_NET=SimpleNamespace(ROT_TYPE="allo_rot6d", TRANS_TYPE="centroid_z", Z_TYPE="REL",
                                        WITH_2D_COORD=True, REGION_ATTENTION=True),
            )
        ),
        TEST=SimpleNamespace(USE_PNP=with_maps, SAVE_RESULTS_ONLY=False, USE_DEPTH_REFINE=with_maps),
        INPUT=SimpleNamespace(WITH_DEPTH=False),
    )


class _Shell(nn.Module):
    """Parameter container reproducing the reference module tree (no forward of its own)."""


def _build_param_tree(root, shapes):
    for key, shape in shapes.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Shell())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))


class GDRN_DoubleMask(nn.Module):
    def __init__(self, cfg, arch="convnext_base", max_batch=64, precision=None):
        """precision: "bf16" (default; tensor-core bf16 operands, the reference's AMP regime) or "bf16x3"
        (split-bf16 GEMMs + fp32 FC stack: reproduces the reference's fp32 forward, see include/gdrn_b200.h).
        None -> cfg.MODEL.POSE_NET.PRECISION if present, else the GDRN_PRECISION environment variable, else bf16."""
        super().__init__()
        net_cfg = cfg.MODEL.POSE_NET
        assert net_cfg.NAME == "GDRN_double_mask", net_cfg.NAME
        self.cfg = cfg
        self.arch = arch
        self.num_classes = int(net_cfg.NUM_CLASSES)
        self.max_batch = max_batch
        if precision is None:
            precision = getattr(net_cfg, "PRECISION", None)
        if precision is None:
            precision = {"0": "bf16", "1": "bf16x3"}.get(os.environ.get("GDRN_PRECISION", "0"), "bf16")
        if precision not in ("bf16", "bf16x3"):
            raise ValueError(f"precision must be 'bf16' or 'bf16x3', got {precision!r}")
        self.precision = precision
        self.neck = None
        if arch not in CONVNEXT_ARCH:
            raise ValueError(f"unknown backbone {arch}")
        pnp = net_cfg.PNP_NET
        if pnp.ROT_TYPE != "allo_rot6d" or pnp.TRANS_TYPE != "centroid_z" or _cfg_get(pnp, "Z_TYPE", "REL") != "REL":
            raise NotImplementedError("the B200 path implements ROT_TYPE=allo_rot6d, TRANS_TYPE=centroid_z, Z_TYPE=REL")
        g = net_cfg.GEO_HEAD
        if not (g.XYZ_CLASS_AWARE and g.MASK_CLASS_AWARE and g.REGION_CLASS_AWARE and g.NUM_REGIONS == 64):
            raise NotImplementedError("the B200 path implements the class-aware, 64-region geometry head")
        shapes = {k: tuple(v.shape) for k, v in make_state_dict(arch, self.num_classes, seed=0).items()}
        _build_param_tree(self, shapes)
        self._handle = None
        self._loaded_version = None
        self._workspace = None
        self._param_version = 0

    # --- weights ---------------------------------------------------------------------------------
    def _load_from_state_dict(self, *args, **kwargs):
        self._param_version += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._param_version += 1
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def mark_weights_dirty(self):
        self._param_version += 1

    def _ensure_engine(self, device):
        L = _lib.lib()
        if self._handle is None:
            h = ctypes.c_void_p()
            _lib.check(L.gdrn_model_create_ex(ctypes.byref(h), self.arch.encode(), self.num_classes, self.max_batch,
                                              1 if self.precision == "bf16x3" else 0), "gdrn_model_create_ex")
            self._handle = h
        if self._loaded_version != self._param_version:
            st = _lib.current_stream()
            for k, v in self.state_dict().items():
                t = v.detach().to(device=device, dtype=torch.float32).contiguous()
                _lib.check(L.gdrn_model_load_tensor(self._handle, k.encode(), _lib.ptr(t), t.numel(), st),
                           f"gdrn_model_load_tensor({k})")
            torch.cuda.current_stream().synchronize()  # sources may be temporaries
            missing = L.gdrn_model_missing(self._handle)
            if missing != 0:
                raise _lib.GdrnError(f"{missing} weight tensors missing after load")
            self._loaded_version = self._param_version

    def _get_workspace(self, B, device):
        need = _lib.lib().gdrn_model_workspace_bytes(self._handle, B)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
        return self._workspace

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().gdrn_model_destroy(self._handle)
        except Exception:
            pass

    # --- forward ---------------------------------------------------------------------------------
    def forward(self, x, gt_xyz=None, gt_xyz_bin=None, gt_mask_trunc=None, gt_mask_visib=None, gt_mask_obj=None,
                gt_mask_full=None, gt_region=None, gt_ego_rot=None, gt_points=None, sym_infos=None, gt_trans=None,
                gt_trans_ratio=None, roi_classes=None, roi_coord_2d=None, roi_coord_2d_rel=None, roi_cams=None,
                roi_centers=None, roi_whs=None, roi_extents=None, resize_ratios=None, do_loss=False,
                return_raw=False):
        if do_loss:
            raise NotImplementedError("training (do_loss=True) is out of scope of the B200 hot path")
        if not x.is_cuda:
            raise _lib.GdrnError("GDRN_DoubleMask.forward needs CUDA tensors (no CPU fallback)")
        for name, t in (("roi_classes", roi_classes), ("roi_coord_2d", roi_coord_2d), ("roi_cams", roi_cams),
                        ("roi_centers", roi_centers), ("roi_whs", roi_whs), ("roi_extents", roi_extents),
                        ("resize_ratios", resize_ratios)):
            if t is None:
                raise ValueError(f"{name} is required")
        dev = x.device
        B = x.shape[0]
        if tuple(x.shape[1:]) != (3, 256, 256):
            raise ValueError(f"roi_img must be [B,3,256,256], got {tuple(x.shape)}")
        self._ensure_engine(dev)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        x = f32(x)
        cls = roi_classes.detach().to(device=dev, dtype=torch.int64).contiguous()
        if roi_cams.dim() == 2:
            roi_cams = roi_cams.unsqueeze(0).expand(B, 3, 3)
        c2d, cams, ctr, whs, ext, rr = (f32(roi_coord_2d), f32(roi_cams), f32(roi_centers), f32(roi_whs),
                                        f32(roi_extents), f32(resize_ratios.reshape(-1)))
        out_rot = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
        out_trans = torch.empty((B, 3), dtype=torch.float32, device=dev)
        out_raw = torch.empty((B, 9), dtype=torch.float32, device=dev) if return_raw else None
        cfg = self.cfg
        want_maps = bool(_cfg_get(cfg, "TEST.USE_PNP", False) or _cfg_get(cfg, "TEST.SAVE_RESULTS_ONLY", False)
                         or _cfg_get(cfg, "TEST.USE_DEPTH_REFINE", False))
        maps = None
        tensors = {}
        if want_maps:
            for name, ch in (("mask", 1), ("full_mask", 1), ("coor_x", 1), ("coor_y", 1), ("coor_z", 1), ("region", 65)):
                tensors[name] = torch.empty((B, ch, 64, 64), dtype=torch.float32, device=dev)
            maps = _lib.GdrnMaps(*[tensors[n].data_ptr() for n in ("mask", "full_mask", "coor_x", "coor_y", "coor_z", "region")])
        ws = self._get_workspace(B, dev)
        L = _lib.lib()
        rc = L.gdrn_model_forward(
            self._handle, _lib.ptr(x), _lib.ptr(cls), _lib.ptr(c2d), _lib.ptr(cams), _lib.ptr(ctr), _lib.ptr(whs),
            _lib.ptr(rr), _lib.ptr(ext), B, _lib.ptr(out_rot), _lib.ptr(out_trans), _lib.ptr(out_raw),
            ctypes.byref(maps) if maps is not None else None, _lib.ptr(ws), ws.numel(), _lib.current_stream())
        _lib.check(rc, "gdrn_model_forward")
        out = {"rot": out_rot, "trans": out_trans}
        out.update(tensors)
        if return_raw:
            out["raw"] = out_raw
        return out

    def capture_graph(self, batch, warmup=2):
        """Capture one forward over the STATIC tensors of ``batch`` (reference data_dict keys: roi_img, roi_classes,
        roi_coord_2d, roi_cams, roi_centers, roi_whs, roi_extents, resize_ratios) into a CUDA graph.

        Returns ``(replay, out_dict)``: ``replay()`` re-runs the ~150 kernel launches of the forward as one graph
        launch on the current stream (inputs are read from the same tensors, so refill them in place);
        ``out_dict`` holds the static output tensors.  CUDA streams and graphs replace a tracing compiler here."""
        kw = dict(roi_classes=batch["roi_classes"], roi_coord_2d=batch["roi_coord_2d"], roi_cams=batch["roi_cams"],
                  roi_centers=batch["roi_centers"], roi_whs=batch["roi_whs"], roi_extents=batch["roi_extents"],
                  resize_ratios=batch["resize_ratios"])
        for k, v in list(kw.items()) + [("roi_img", batch["roi_img"])]:
            if not (v.is_cuda and v.is_contiguous()):
                raise _lib.GdrnError(f"capture_graph: {k} must be a contiguous CUDA tensor")
        if batch["roi_img"].dtype != torch.float32 or batch["roi_classes"].dtype != torch.int64:
            raise _lib.GdrnError("capture_graph: roi_img must be float32 and roi_classes int64 (no hidden copies)")
        side = torch.cuda.Stream(device=batch["roi_img"].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.forward(batch["roi_img"], **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward(batch["roi_img"], **kw)
        return graph.replay, out

    def debug_read(self, name, B, numel):
        dst = torch.empty(numel, dtype=torch.float32, device=self._workspace.device)
        n = _lib.lib().gdrn_model_debug_read(self._handle, name.encode(), B, _lib.ptr(dst), _lib.ptr(self._workspace),
                                             _lib.current_stream())
        if n < 0:
            raise _lib.GdrnError(f"debug_read({name}) failed: {_lib.last_error()}")
        return dst[:n]


def build_model_optimizer(cfg, is_test=True):
    """Reference entry point (GDRN_double_mask.py:539-615).  Only the test-time build is in scope."""
    if not is_test:
        raise NotImplementedError("training is out of scope of the B200 hot path (SURVEY.md §2 row 17/19)")
    bb_type = _cfg_get(cfg, "MODEL.POSE_NET.BACKBONE.INIT_CFG.type", "timm/convnext_base")
    arch = bb_type.split("/")[-1]
    model = GDRN_DoubleMask(cfg, arch=arch)
    return model, None

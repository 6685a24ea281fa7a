"""GdrnPredictor -- the single-image API of the reference (core/gdrn_modeling/demo/predictor_gdrn.py:44-476),
routed to the B200 hot path.

Same method names and data contracts: ``preprocessing(outputs, image, depth_img) -> data_dict``,
``inference(data_dict) -> out_dict``, ``postprocessing(data_dict, out_dict) -> {obj_name: 4x4 pose}``,
``process_depth_refine(data_dict, out_dict)``.  Differences that the absence of datasets/checkpoints forces:
the constructor takes the already-loaded pieces (state_dict or checkpoint path, camera matrix, meshes, extents)
instead of dataset paths.  ``inference`` and ``process_depth_refine`` run entirely in libgdrn_b200.so;
``preprocessing`` (crop_resize_by_warp_affine, SURVEY.md §8f rank 1, "next") is a plain torch bilinear crop for
now and is NOT part of the measured hot path.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .gdrn_model import GDRN_DoubleMask, default_cfg
from .renderer import Model3D, depth_refine, get_K_crop_resize


class GdrnPredictor:
    def __init__(self, cam, objs, extents, models=None, state_dict=None, ckpt_file_path=None, num_classes=None,
                 use_depth_refine=False, depth_refine_iter=2, depth_refine_threshold=0.8, depth_scale=1.0,
                 dzi_pad_scale=1.5, device="cuda", cfg=None):
        """cam: 3x3 intrinsics; objs: {obj_id: name}; extents: {obj_id: (3,)} metres;
        models: {obj_id: (verts[V,3] metres, faces[F,3])} (needed for depth refine)."""
        self.cam = np.asarray(cam, np.float32)
        self.objs = dict(objs)
        self.cls_names = list(self.objs.values())
        self.obj_ids = list(self.objs.keys())
        self.extents = {k: np.asarray(v, np.float32) for k, v in extents.items()}
        self.depth_scale = depth_scale
        self.device = torch.device(device)
        nc = num_classes or len(self.obj_ids)
        self.cfg = cfg or default_cfg(num_classes=nc, with_maps=use_depth_refine)
        self.cfg.TEST.USE_PNP = False
        self.cfg.TEST.USE_DEPTH_REFINE = bool(use_depth_refine)
        self.cfg.TEST.DEPTH_REFINE_ITER = depth_refine_iter
        self.cfg.TEST.DEPTH_REFINE_THRESHOLD = depth_refine_threshold
        self.dzi_pad_scale = dzi_pad_scale
        self.model = GDRN_DoubleMask(self.cfg)
        if state_dict is None and ckpt_file_path is not None:
            ck = torch.load(ckpt_file_path, map_location="cpu")
            state_dict = ck.get("model", ck)
            state_dict = {k[len("_module."):] if k.startswith("_module.") else k: v for k, v in state_dict.items()}
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.model.to(self.device)
        self.ren_models = None
        if models is not None:
            self.ren_models = [Model3D(*models[i], device=self.device) for i in self.obj_ids]

    # ---- preprocessing (predictor_gdrn.py:301-476; bilinear affine crop) -----------------------
    def preprocessing(self, outputs, image, depth_img=None):
        """outputs: [n,7] detections (x1,y1,x2,y2,score,cls_score,cls) like the YOLOX stage; image: HxWx3 BGR uint8."""
        dev = self.device
        det = torch.as_tensor(outputs, dtype=torch.float32)
        n = det.shape[0]
        im = torch.as_tensor(image).to(dev).permute(2, 0, 1).float()[None] / 255.0
        H, W = im.shape[-2:]
        x1, y1, x2, y2 = det[:, 0], det[:, 1], det[:, 2], det[:, 3]
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        bw, bh = (x2 - x1).clamp_min(1), (y2 - y1).clamp_min(1)
        scale = torch.clamp(torch.maximum(bw, bh) * self.dzi_pad_scale, max=float(max(H, W)))

        def crop(src, res):
            u = torch.arange(res, dtype=torch.float32)
            sx = cx[:, None] + (u[None] - res / 2) * (scale[:, None] / res)
            sy = cy[:, None] + (u[None] - res / 2) * (scale[:, None] / res)
            gx = (sx / (W - 1) * 2 - 1)[:, None, :].expand(n, res, res)
            gy = (sy / (H - 1) * 2 - 1)[:, :, None].expand(n, res, res)
            grid = torch.stack([gx, gy], -1).to(dev)
            return F.grid_sample(src.expand(n, -1, -1, -1), grid, mode="bilinear", padding_mode="zeros", align_corners=True)

        roi_img = crop(im, 256)
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32) / H, torch.arange(W, dtype=torch.float32) / W, indexing="ij")
        coord = torch.stack([xx, yy])[None].to(dev)
        roi_coord_2d = crop(coord, 64)
        cls = det[:, 6].long()
        ext = torch.stack([torch.from_numpy(self.extents[self.obj_ids[int(c)]]) for c in cls])
        data = {
            "roi_img": roi_img, "roi_cls": cls.to(dev), "roi_coord_2d": roi_coord_2d,
            "roi_cam": torch.from_numpy(self.cam)[None].repeat(n, 1, 1).to(dev),
            "roi_center": torch.stack([cx, cy], 1).to(dev), "roi_wh": torch.stack([bw, bh], 1).to(dev),
            "scale": scale.to(dev), "resize_ratio": (64.0 / scale).to(dev), "roi_extent": ext.to(dev),
            "score": det[:, 4] * det[:, 5] if det.shape[1] > 5 else det[:, 4], "bbox_est": det[:, :4],
        }
        if depth_img is not None:
            d = torch.as_tensor(np.asarray(depth_img, np.float32) * self.depth_scale).to(dev)[None, None]
            data["roi_depth"] = crop(d, 64)[:, 0]
        return data

    # ---- inference (predictor_gdrn.py:122-147) --------------------------------------------------
    @torch.no_grad()
    def inference(self, data_dict):
        out = self.model(
            data_dict["roi_img"], roi_classes=data_dict["roi_cls"], roi_cams=data_dict["roi_cam"],
            roi_whs=data_dict["roi_wh"], roi_centers=data_dict["roi_center"], resize_ratios=data_dict["resize_ratio"],
            roi_coord_2d=data_dict.get("roi_coord_2d"), roi_extents=data_dict.get("roi_extent"))
        torch.cuda.synchronize()
        return out

    # ---- postprocessing (predictor_gdrn.py:149-191) ---------------------------------------------
    def postprocessing(self, data_dict, out_dict):
        rot = out_dict["rot"]
        trans = out_dict["trans"]
        if self.cfg.TEST.USE_DEPTH_REFINE:
            trans = self.process_depth_refine(data_dict, out_dict)
        R = rot.detach().cpu().numpy()
        t = trans.detach().cpu().numpy()
        data_dict["cur_res"] = []
        poses = {}
        for i in range(R.shape[0]):
            oid = self.obj_ids[int(data_dict["roi_cls"][i])]
            data_dict["cur_res"].append({"obj_id": oid, "score": float(data_dict["score"][i]),
                                         "bbox_est": np.asarray(data_dict["bbox_est"][i]), "R": R[i], "t": t[i]})
            pose = np.eye(4)
            pose[:3, :3], pose[:3, 3] = R[i], t[i]
            poses[self.objs.get(oid)] = pose
        return poses

    # ---- fast depth refine (predictor_gdrn.py:195-286), batched on the GPU ----------------------
    def process_depth_refine(self, inputs, out_dict):
        if self.ren_models is None:
            raise RuntimeError("depth refine needs object meshes (models=...)")
        n = out_dict["rot"].shape[0]
        crop_xy = inputs["roi_center"] - inputs["scale"].view(n, 1) / 2
        K_crop = get_K_crop_resize(inputs["roi_cam"], crop_xy, (64.0 / inputs["scale"]).view(n, 1))
        xyz = torch.cat([out_dict["coor_x"], out_dict["coor_y"], out_dict["coor_z"]], dim=1)
        return depth_refine([m.vertices for m in self.ren_models], [m.faces for m in self.ren_models], out_dict["rot"],
                            out_dict["trans"], K_crop, xyz, out_dict["mask"], inputs["roi_depth"],
                            iters=self.cfg.TEST.DEPTH_REFINE_ITER, thresh=self.cfg.TEST.DEPTH_REFINE_THRESHOLD,
                            mesh_ids=inputs["roi_cls"])

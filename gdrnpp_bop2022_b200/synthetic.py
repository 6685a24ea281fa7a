"""Seeded synthetic weights and ROI batches.

There is no network for datasets or checkpoints, so both sides of every parity
test (CUDA path and oracle) consume the ``state_dict`` and the batch made here.

Weight distribution: the reference's own random init is numerically degenerate
(ConvNeXt layer-scale gamma = 1e-6, head convs N(0, 1e-3); SURVEY.md §7 "hard
parts"), which would make R|t parity pure noise.  We use a variance-preserving
init instead (He-scaled convs/linears, gamma ~ U(0.1, 0.4), norm affine params
perturbed around identity) and record it here; key names and shapes are exactly
those of a reference checkpoint.
"""
import math

import numpy as np
import torch

CONVNEXT_ARCH = {
    "convnext_tiny": ((3, 3, 9, 3), (96, 192, 384, 768)),
    "convnext_small": ((3, 3, 27, 3), (96, 192, 384, 768)),
    "convnext_base": ((3, 3, 27, 3), (128, 256, 512, 1024)),
}

YCBV_K = ((1066.778, 0.0, 312.9869), (0.0, 1067.487, 241.3109), (0.0, 0.0, 1.0))  # ref/ycbv.py:89


def state_dict_shapes(arch="convnext_base", num_classes=21, num_regions=64):
    """{key: shape} of a reference checkpoint without materialising any weights."""
    depths, dims = CONVNEXT_ARCH[arch]
    sh = {}

    def norm(prefix, c):
        sh[prefix + ".weight"] = (c,)
        sh[prefix + ".bias"] = (c,)

    b = "backbone."
    sh[b + "stem_0.weight"] = (dims[0], 3, 4, 4)
    sh[b + "stem_0.bias"] = (dims[0],)
    norm(b + "stem_1", dims[0])
    for s in range(4):
        C = dims[s]
        if s > 0:
            norm(b + f"stages_{s}.downsample.0", dims[s - 1])
            sh[b + f"stages_{s}.downsample.1.weight"] = (C, dims[s - 1], 2, 2)
            sh[b + f"stages_{s}.downsample.1.bias"] = (C,)
        for i in range(depths[s]):
            p = b + f"stages_{s}.blocks.{i}."
            sh[p + "conv_dw.weight"] = (C, 1, 7, 7)
            sh[p + "conv_dw.bias"] = (C,)
            norm(p + "norm", C)
            sh[p + "mlp.fc1.weight"] = (4 * C, C)
            sh[p + "mlp.fc1.bias"] = (4 * C,)
            sh[p + "mlp.fc2.weight"] = (C, 4 * C)
            sh[p + "mlp.fc2.bias"] = (C,)
            sh[p + "gamma"] = (C,)
    h = "geo_head_net."
    sh[h + "features.0.weight"] = (dims[3], 256, 3, 3)
    norm(h + "features.1", 256)
    for i in (3, 4, 6, 7, 9, 10):
        sh[h + f"features.{i}.conv.weight"] = (256, 256, 3, 3)
        norm(h + f"features.{i}.gn", 256)
    out_dim = num_classes * (2 + 3 + num_regions + 1)
    sh[h + "out_layer.weight"] = (out_dim, 256, 1, 1)
    sh[h + "out_layer.bias"] = (out_dim,)
    p = "pnp_net."
    n_in = 3 + 2 + num_regions
    for j, i in enumerate((0, 3, 6)):
        sh[p + f"features.{i}.weight"] = (128, n_in if j == 0 else 128, 3, 3)
        norm(p + f"features.{i+1}", 128)
    sh[p + "fc1.weight"] = (1024, 8192)
    sh[p + "fc1.bias"] = (1024,)
    sh[p + "fc2.weight"] = (256, 1024)
    sh[p + "fc2.bias"] = (256,)
    sh[p + "fc_r.weight"] = (6, 256)
    sh[p + "fc_r.bias"] = (6,)
    sh[p + "fc_t.weight"] = (3, 256)
    sh[p + "fc_t.bias"] = (3,)
    return sh


def make_state_dict(arch="convnext_base", num_classes=21, num_regions=64, seed=0):
    """fp32 state_dict with reference key names: backbone.* / geo_head_net.* / pnp_net.*"""
    g = torch.Generator().manual_seed(seed)
    depths, dims = CONVNEXT_ARCH[arch]
    sd = {}

    def randn(shape, std):
        return torch.randn(shape, generator=g) * std

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    def norm_affine(prefix, c):
        sd[prefix + ".weight"] = uni((c,), 0.8, 1.2)
        sd[prefix + ".bias"] = randn((c,), 0.05)

    b = "backbone."
    sd[b + "stem_0.weight"] = randn((dims[0], 3, 4, 4), 1.0 / math.sqrt(48) * 2.0)
    sd[b + "stem_0.bias"] = randn((dims[0],), 0.1)
    norm_affine(b + "stem_1", dims[0])
    for s in range(4):
        C = dims[s]
        if s > 0:
            norm_affine(b + f"stages_{s}.downsample.0", dims[s - 1])
            sd[b + f"stages_{s}.downsample.1.weight"] = randn((C, dims[s - 1], 2, 2), 1.0 / math.sqrt(4 * dims[s - 1]))
            sd[b + f"stages_{s}.downsample.1.bias"] = randn((C,), 0.05)
        for i in range(depths[s]):
            p = b + f"stages_{s}.blocks.{i}."
            sd[p + "conv_dw.weight"] = randn((C, 1, 7, 7), 1.0 / 7.0)
            sd[p + "conv_dw.bias"] = randn((C,), 0.05)
            norm_affine(p + "norm", C)
            sd[p + "mlp.fc1.weight"] = randn((4 * C, C), 1.0 / math.sqrt(C))
            sd[p + "mlp.fc1.bias"] = randn((4 * C,), 0.05)
            sd[p + "mlp.fc2.weight"] = randn((C, 4 * C), math.sqrt(2.0) / math.sqrt(4 * C))
            sd[p + "mlp.fc2.bias"] = randn((C,), 0.05)
            sd[p + "gamma"] = uni((C,), 0.1, 0.4)

    h = "geo_head_net."
    in_dim = dims[3]
    sd[h + "features.0.weight"] = randn((in_dim, 256, 3, 3), math.sqrt(2.0) / math.sqrt(in_dim * 9 / 4))
    norm_affine(h + "features.1", 256)
    for i in (3, 4, 6, 7, 9, 10):
        sd[h + f"features.{i}.conv.weight"] = randn((256, 256, 3, 3), math.sqrt(2.0) / math.sqrt(256 * 9))
        norm_affine(h + f"features.{i}.gn", 256)
    out_dim = num_classes * (2 + 3 + num_regions + 1)
    sd[h + "out_layer.weight"] = randn((out_dim, 256, 1, 1), 1.0 / math.sqrt(256))
    sd[h + "out_layer.bias"] = randn((out_dim,), 0.1)
    # xyz outputs are expected in [0,1] by the PnP denormalisation: centre them at 0.5
    sd[h + "out_layer.bias"][2 * num_classes : 5 * num_classes] += 0.5
    sd[h + "out_layer.weight"][2 * num_classes : 5 * num_classes] *= 0.25

    p = "pnp_net."
    n_in = 3 + 2 + num_regions
    for j, i in enumerate((0, 3, 6)):
        cin = n_in if j == 0 else 128
        sd[p + f"features.{i}.weight"] = randn((128, cin, 3, 3), math.sqrt(2.0) / math.sqrt(cin * 9))
        norm_affine(p + f"features.{i+1}", 128)
    sd[p + "fc1.weight"] = randn((1024, 8192), math.sqrt(2.0) / math.sqrt(8192))
    sd[p + "fc1.bias"] = randn((1024,), 0.05)
    sd[p + "fc2.weight"] = randn((256, 1024), math.sqrt(2.0) / math.sqrt(1024))
    sd[p + "fc2.bias"] = randn((256,), 0.05)
    sd[p + "fc_r.weight"] = randn((6, 256), 1.0 / math.sqrt(256))
    sd[p + "fc_r.bias"] = randn((6,), 0.3)
    sd[p + "fc_t.weight"] = randn((3, 256), 0.2 / math.sqrt(256))
    sd[p + "fc_t.bias"] = torch.tensor([0.0, 0.0, 1.0]) + randn((3,), 0.02)
    return sd


def make_batch(B=64, seed=0, num_classes=21, im_w=640, im_h=480, in_res=256, out_res=64):
    """Synthetic ROI batch with the tensor contract of datasets/data_loader.py:647-818
    (roi_img in [0,1) because PIXEL_MEAN 0 / PIXEL_STD 255)."""
    g = torch.Generator().manual_seed(1000 + seed)
    roi_img = torch.rand((B, 3, in_res, in_res), generator=g)
    roi_classes = torch.randint(0, num_classes, (B,), generator=g)
    cx = torch.rand((B,), generator=g) * im_w
    cy = torch.rand((B,), generator=g) * im_h
    bw = torch.rand((B,), generator=g) * 160 + 40
    bh = torch.rand((B,), generator=g) * 160 + 40
    scale = torch.clamp(torch.maximum(bw, bh) * 1.5, max=float(max(im_w, im_h)))
    resize_ratios = out_res / scale
    u = torch.arange(out_res, dtype=torch.float32)
    # affine crop: dst pixel u <-> src = c + (u - out_res/2) * scale/out_res (data_utils.py:136-184)
    sx = cx[:, None] + (u[None, :] - out_res / 2) * (scale[:, None] / out_res)
    sy = cy[:, None] + (u[None, :] - out_res / 2) * (scale[:, None] / out_res)
    gx = torch.where((sx >= 0) & (sx <= im_w - 1), sx / im_w, torch.zeros_like(sx))
    gy = torch.where((sy >= 0) & (sy <= im_h - 1), sy / im_h, torch.zeros_like(sy))
    coord = torch.stack(
        [gx[:, None, :].expand(B, out_res, out_res), gy[:, :, None].expand(B, out_res, out_res)], dim=1
    ).contiguous()
    roi_cams = torch.tensor(YCBV_K, dtype=torch.float32)[None].repeat(B, 1, 1)
    extents = torch.rand((B, 3), generator=g) * 0.2 + 0.05
    return {
        "roi_img": roi_img,
        "roi_classes": roi_classes,
        "roi_coord_2d": coord,
        "roi_cams": roi_cams,
        "roi_centers": torch.stack([cx, cy], dim=1),
        "roi_whs": torch.stack([bw, bh], dim=1),
        "resize_ratios": resize_ratios,
        "roi_extents": extents,
    }


def make_icosphere_mesh(subdiv=3, extent=(0.1, 0.1, 0.1)):
    """Unit icosphere scaled to ``extent`` (full widths, metres): verts [V,3] f32, faces [F,3] i32."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    verts = np.array(v, dtype=np.float64) * (np.array(extent, dtype=np.float64) / 2.0)
    return verts.astype(np.float32), np.array(f, dtype=np.int32)

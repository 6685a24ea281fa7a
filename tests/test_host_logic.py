"""CPU tests of the host-side logic: synthetic generators, the reference-shaped model surface, error behaviour
without a GPU, ROI sharding and the world-size-2 pose all-gather over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_state_dict_keys_and_shapes():
    from gdrnpp_bop2022_b200.synthetic import make_state_dict

    sd = make_state_dict()
    assert sum(v.numel() for v in sd.values()) == 102873543          # 102.87 M params
    assert sd["backbone.stages_2.blocks.26.mlp.fc1.weight"].shape == (2048, 512)
    assert sd["geo_head_net.features.0.weight"].shape == (1024, 256, 3, 3)
    assert sd["geo_head_net.out_layer.weight"].shape == (1470, 256, 1, 1)
    assert sd["pnp_net.features.0.weight"].shape == (128, 69, 3, 3)
    assert sd["pnp_net.fc1.weight"].shape == (1024, 8192)
    sd2 = make_state_dict()
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)               # seeded


def test_make_batch_contract():
    from gdrnpp_bop2022_b200.synthetic import make_batch

    b = make_batch(B=5, seed=1)
    assert b["roi_img"].shape == (5, 3, 256, 256) and 0 <= b["roi_img"].min() and b["roi_img"].max() < 1
    assert b["roi_classes"].dtype == torch.int64 and b["roi_classes"].max() < 21
    assert b["roi_coord_2d"].shape == (5, 2, 64, 64) and b["roi_cams"].shape == (5, 3, 3)
    assert torch.allclose(b["resize_ratios"] * torch.clamp(torch.maximum(b["roi_whs"][:, 0], b["roi_whs"][:, 1]) * 1.5, max=640.0),
                          torch.full((5,), 64.0))


def test_model_surface_matches_reference_checkpoint_layout():
    from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, build_model_optimizer, default_cfg
    from gdrnpp_bop2022_b200.synthetic import make_state_dict

    model, opt = build_model_optimizer(default_cfg(), is_test=True)
    assert opt is None and isinstance(model, GDRN_DoubleMask) and model.neck is None
    sd = make_state_dict()
    assert set(model.state_dict().keys()) == set(sd.keys())
    missing = model.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert torch.equal(model.backbone.stages_1.downsample[1].weight if False else model.state_dict()["backbone.stages_1.downsample.1.weight"],
                       sd["backbone.stages_1.downsample.1.weight"])
    assert hasattr(model, "backbone") and hasattr(model, "geo_head_net") and hasattr(model, "pnp_net")
    with pytest.raises(NotImplementedError):
        build_model_optimizer(default_cfg(), is_test=False)


def test_forward_without_gpu_fails_loudly():
    from gdrnpp_bop2022_b200 import _lib
    from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, default_cfg

    m = GDRN_DoubleMask(default_cfg())
    with pytest.raises(_lib.GdrnError):
        m(torch.zeros(1, 3, 256, 256), roi_classes=torch.zeros(1, dtype=torch.long), roi_coord_2d=torch.zeros(1, 2, 64, 64),
          roi_cams=torch.eye(3)[None], roi_centers=torch.zeros(1, 2), roi_whs=torch.ones(1, 2), roi_extents=torch.ones(1, 3),
          resize_ratios=torch.ones(1))
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 256, 256), do_loss=True)


def test_native_op_wrappers_reject_cpu_tensors():
    from gdrnpp_bop2022_b200 import native_ops

    with pytest.raises(RuntimeError, match="CUDA"):
        native_ops.ransac_voting.generate_hypothesis(torch.zeros(4, 2, 2), torch.zeros(4, 2), torch.zeros(3, 2, 2, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA"):
        native_ops.nnd(torch.zeros(1, 4, 3), torch.zeros(1, 5, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        native_ops.flow_cuda.forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), torch.zeros(1, 3, 4), torch.zeros(1, 3, 3))


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gdrnpp_bop2022_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in src and "import oracle" not in src and "oracle/" not in src, f


def test_shard_range_and_pose_packing():
    from gdrnpp_bop2022_b200.dist import pack_poses, shard_range, unpack_poses

    for total, world in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        covered = []
        for r in range(world):
            b, e = shard_range(total, r, world)
            assert 0 <= b <= e <= total
            covered += list(range(b, e))
        assert covered == list(range(total))
    assert shard_range(4096, 3, 8) == (1536, 2048)
    R, t = torch.randn(5, 3, 3), torch.randn(5, 3)
    R2, t2 = unpack_poses(pack_poses(R, t))
    assert torch.equal(R, R2) and torch.equal(t, t2)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gdrnpp_bop2022_b200.dist import all_gather_poses, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = int(sys.argv[2])
g = torch.Generator().manual_seed(0)
R, t = torch.randn(total, 3, 3, generator=g), torch.randn(total, 3, generator=g)
b, e = shard_range(total, rank, world)
for tot in (total, None):
    Rg, tg = all_gather_poses(R[b:e].clone(), t[b:e].clone(), total=tot)
    assert torch.equal(Rg, R) and torch.equal(tg, t), (rank, tot)
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("total", [64, 7, 1])
def test_all_gather_poses_gloo_world2(tmp_path, total):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29000 + total), str(script), ROOT, str(total)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_state_dict_shapes_match_generated_weights():
    """GDRN_DoubleMask builds its parameter tree from state_dict_shapes() (no 103 M random numbers per constructor);
    the table must stay in lock-step with the generator the tests and the bench load."""
    from gdrnpp_bop2022_b200.synthetic import make_state_dict, state_dict_shapes

    for arch in ("convnext_base", "convnext_tiny"):
        sh, sd = state_dict_shapes(arch), make_state_dict(arch)
        assert list(sh.keys()) == list(sd.keys())
        assert all(tuple(sd[k].shape) == tuple(sh[k]) for k in sh)


def test_ply_roundtrip_and_reference_semantics(tmp_path):
    """load_ply mirrors lib/pysixd/inout.py:489 (pts scaled by vertex_scale, triangular faces) for the ascii and the
    binary_little_endian encodings BOP models ship in."""
    from gdrnpp_bop2022_b200.ply import load_ply, save_ply
    from gdrnpp_bop2022_b200.synthetic import make_icosphere_mesh

    v, f = make_icosphere_mesh(2, (120.0, 80.0, 100.0))   # millimetres, like BOP models
    for binary in (True, False):
        p = str(tmp_path / ("m_%d.ply" % binary))
        save_ply(p, v, f, binary=binary)
        m = load_ply(p, vertex_scale=0.001)
        assert m["pts"].shape == v.shape and m["faces"].shape == f.shape
        assert np.allclose(m["pts"], v.astype(np.float64) * 0.001, atol=1e-9 if binary else 1e-6)
        assert np.array_equal(m["faces"], f)
    # extra vertex properties (normals, colours) and a header comment, ascii
    p = str(tmp_path / "n.ply")
    with open(p, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n"
                 "0 0 0 0 0 1 255 0 0\n1 0 0 0 0 1 0 255 0\n0 1 0 0 0 1 0 0 255\n3 0 1 2\n")
    m = load_ply(p)
    assert m["pts"].shape == (3, 3) and m["normals"][0, 2] == 1 and m["colors"][1, 1] == 255 and m["faces"].tolist() == [[0, 1, 2]]


def test_reference_style_config_loader(tmp_path):
    """load_py_config: module-level dicts + `_base_` inheritance with key-wise merge, like mmcv.Config.fromfile."""
    from gdrnpp_bop2022_b200.gdrn_model import GDRN_DoubleMask, load_py_config

    (tmp_path / "base.py").write_text(
        "MODEL = dict(PIXEL_STD=[255.0, 255.0, 255.0], POSE_NET=dict(NAME='GDRN_double_mask', NUM_CLASSES=13, OUTPUT_RES=64,\n"
        "    BACKBONE=dict(INIT_CFG=dict(type='timm/resnet34')),\n"
        "    GEO_HEAD=dict(NUM_REGIONS=64, XYZ_CLASS_AWARE=False, MASK_CLASS_AWARE=False, REGION_CLASS_AWARE=False, MASK_THR_TEST=0.5),\n"
        "    PNP_NET=dict(ROT_TYPE='allo_rot6d', TRANS_TYPE='centroid_z', Z_TYPE='REL')))\n"
        "TEST = dict(USE_PNP=False, USE_DEPTH_REFINE=False, DEPTH_REFINE_ITER=2)\nINPUT = dict(DZI_PAD_SCALE=1.5)\n")
    (tmp_path / "ycbv.py").write_text(
        "_base_ = ['base.py']\n"
        "MODEL = dict(POSE_NET=dict(NUM_CLASSES=21, BACKBONE=dict(INIT_CFG=dict(type='timm/convnext_base')),\n"
        "    GEO_HEAD=dict(XYZ_CLASS_AWARE=True, MASK_CLASS_AWARE=True, REGION_CLASS_AWARE=True)))\n")
    cfg = load_py_config(str(tmp_path / "ycbv.py"))
    pn = cfg.MODEL.POSE_NET
    assert pn.NUM_CLASSES == 21 and pn.BACKBONE.INIT_CFG.type == "timm/convnext_base" and pn.GEO_HEAD.NUM_REGIONS == 64
    assert pn.GEO_HEAD.XYZ_CLASS_AWARE is True and cfg.TEST.DEPTH_REFINE_ITER == 2 and cfg.MODEL.PIXEL_STD[0] == 255.0
    m = GDRN_DoubleMask(cfg)          # the loaded config drives the model surface like the reference's mmcv Config
    assert m.num_classes == 21 and m.precision == "bf16x3"


def test_b_inv_matches_the_reference_contract():
    """ransac_voting_gpu.py:107-120: batched inverse; a singular batch yields the identity instead of raising."""
    import torch

    from gdrnpp_bop2022_b200.native_ops import b_inv

    m = torch.randn(6, 3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) + 3 * torch.eye(3, dtype=torch.float64)
    assert (b_inv(m) @ m - torch.eye(3, dtype=torch.float64)).abs().max().item() < 1e-12
    assert torch.equal(b_inv(torch.zeros(2, 2, 2)), torch.eye(2).expand(2, 2, 2))


def test_covariance_weights_follow_the_reference_loop():
    """un_pnp_utils.py:96-104 restated as the reference's per-point loop (np.max(np.linalg.eigvals(C)), 0 for C[0,0] < 1e-5)."""
    import numpy as np

    from gdrnpp_bop2022_b200.native_ops import covariance_weights

    rs = np.random.RandomState(4)
    A = rs.randn(40, 2, 2)
    covars = A @ A.transpose(0, 2, 1) * rs.uniform(0.01, 4.0, size=(40, 1, 1))
    covars[5] = 0.0
    covars[17, 0, 0] = 1e-6
    ref = []
    for pi in range(40):
        if covars[pi, 0, 0] < 1e-5:
            ref.append(0.0)
        else:
            ref.append(1.0 / np.max(np.linalg.eigvals(covars[pi])))
    got = covariance_weights(covars)
    assert got.dtype == np.float64 and np.allclose(got, np.asarray(ref), rtol=1e-13, atol=0)
    assert got[5] == 0.0 and got[17] == 0.0

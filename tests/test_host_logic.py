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

"""CPU tests of the host-side logic: the module mirror over the oracle ops, config, sharding (gloo, 2 ranks)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT
from oracle import torch_ops
from pdanet_b200.config import load_config
from pdanet_b200.iassd import build_model
from pdanet_b200.runner import shard_scenes
from pdanet_b200.synthetic import make_batch


def test_parameter_count_matches_the_reference_model():
    """6.37 M parameters for PDA-SSD KITTI (analytic from the reference's module constructors)."""
    m = build_model(load_config("kitti"))
    assert sum(p.numel() for p in m.parameters()) == 6369398
    keys = m.state_dict().keys()
    for k in ["backbone_3d.SA_modules.0.mlps.0.0.weight", "backbone_3d.SA_modules.1.Local_pointformer.0.self_attn.in_proj_weight",
              "backbone_3d.SA_modules.1.point_density.0.densitynet.mlp_convs.0.weight", "backbone_3d.SA_modules.4.ctr_reg.weight",
              "backbone_3d.SA_modules.5.aggregation_layer.0.weight", "point_head.cls_center_layers.6.weight", "global_step"]:
        assert k in keys, k


def test_cpu_forward_shapes_small_scene():
    cfg = load_config("kitti")
    torch.manual_seed(0)
    m = build_model(cfg, ops=torch_ops, nms_utils=torch_ops.nms_utils, batched_post_processing=False).eval()
    with torch.no_grad():
        bd = m.backbone_3d(make_batch(2, 16384, cfg.POINT_CLOUD_RANGE))
    assert [tuple(x.shape) for x in bd["encoder_xyz"]] == [(2, 16384, 3), (2, 4096, 3), (2, 1024, 3), (2, 512, 3),
                                                           (2, 256, 3), (2, 256, 3), (2, 256, 3)]
    assert bd["centers_features"].shape == (512, 512)
    assert bd["sa_ins_preds"][1].shape == (2, 1024, 4) and bd["sa_ins_preds"][0] == []


def test_product_ops_refuse_cpu_tensors():
    """No CPU fallback: the product namespace raises instead of computing on the host."""
    from pdanet_b200 import pointnet2_utils as ops
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(torch.rand(1, 64, 3), 8)
    with pytest.raises(RuntimeError):
        ops.ball_query(0.5, 4, torch.rand(1, 64, 3), torch.rand(1, 8, 3))
    with pytest.raises(RuntimeError):
        ops.topk_ctr_sample(torch.rand(1, 64, 3), 8)


def test_shard_scenes_partitions():
    for n, w in [(16, 1), (16, 2), (16, 8), (5, 2), (3, 4)]:
        ids = [i for r in range(w) for i in shard_scenes(n, w, r)]
        assert ids == list(range(n))


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import torch_ops
from pdanet_b200.config import load_config
from pdanet_b200.iassd import build_model
from pdanet_b200.runner import shard_scenes
from pdanet_b200.synthetic import make_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = load_config("kitti")
torch.manual_seed(0)
m = build_model(cfg, ops=torch_ops, nms_utils=torch_ops.nms_utils, batched_post_processing=False).eval()
N, S = 4096, 4
ids = shard_scenes(S, world, rank)
batch = make_batch(len(ids), N, cfg.POINT_CLOUD_RANGE, first_scene=ids.start)
with torch.no_grad():
    preds, _ = m(batch)
sig = torch.tensor([[p["pred_boxes"].shape[0], float(p["pred_boxes"].double().sum())] for p in preds], dtype=torch.float64)
gathered = [torch.zeros_like(sig) for _ in range(world)]
dist.all_gather(gathered, sig)            # result collection only; the data path itself has no collective
if rank == 0:
    with torch.no_grad():
        whole, _ = m(make_batch(S, N, cfg.POINT_CLOUD_RANGE))
    want = torch.tensor([[p["pred_boxes"].shape[0], float(p["pred_boxes"].double().sum())] for p in whole], dtype=torch.float64)
    got = torch.cat(gathered)
    assert torch.allclose(got, want, rtol=1e-9), (got, want)
    print("SHARD_OK")
dist.destroy_process_group()
"""


def test_two_rank_scene_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), str(ROOT)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout


def _forced_ops(base, picks):
    import types
    ns = types.SimpleNamespace(**{k: getattr(base, k) for k in dir(base) if not k.startswith("_")})
    queue = list(picks)
    ns.topk_ctr_sample = lambda cls_features, npoint: queue.pop(0).to(cls_features.device).contiguous()
    return ns


def replay_reference_backbone_golden(ops_base, device, rtol):
    """Our backbone vs tests/golden/ref_backbone_kitti.npz — outputs of the REFERENCE's own IASSD_Backbone /
    pointnet2_modules.py run over the oracle ops (tests/golden/make_module_golden.py).  Weights come from the same
    seed (the generator asserts the seeded state_dicts are identical); the class-aware layers replay the reference's
    picks because torch.topk's tie order is unspecified."""
    import numpy as np
    from conftest import GOLDEN
    from pdanet_b200.iassd_backbone import IASSD_Backbone
    z = np.load(GOLDEN / "ref_backbone_kitti.npz")
    cfg = load_config("kitti")
    torch.manual_seed(int(z["seed"]))
    bb = IASSD_Backbone(cfg.MODEL.BACKBONE_3D, num_class=3, input_channels=4, ops=ops_base).eval().to(device)
    forced = _forced_ops(ops_base, [torch.from_numpy(z["sample_idx_L2"]), torch.from_numpy(z["sample_idx_L3"])])
    for mod in bb.SA_modules:
        if hasattr(mod, "ops"):
            mod.ops = forced
    batch = make_batch(int(z["batch"]), int(z["npoints"]), cfg.POINT_CLOUD_RANGE, duplicate_frac=float(z["duplicate_frac"]))
    with torch.no_grad():
        out = bb({"batch_size": batch["batch_size"], "points": batch["points"].to(device)})

    def close(got, want, what):
        got, want = got.detach().cpu(), torch.from_numpy(want)
        err = (got - want).abs().max().item()
        assert err <= rtol * (want.abs().max().item() + 1e-12), f"{what}: {err:.3e}"

    for lvl in (0, 1):  # D-FPS layers: exact indices
        assert torch.equal(out["sample_list_id"][lvl].cpu(), torch.from_numpy(z[f"sample_idx_L{lvl}"])), f"FPS L{lvl}"
    close(out["centers"], z["centers"], "centers")
    close(out["centers_origin"], z["centers_origin"], "centers_origin")
    close(out["ctr_offsets"][:, 1:], z["ctr_offsets"][:, 1:], "ctr_offsets")
    close(out["centers_features"][::4, ::8], z["centers_features_strided"], "centers_features")
    close(out["sa_ins_preds"][1][:, ::8], z["cls_L1_strided"], "cls L1")
    close(out["sa_ins_preds"][2][:, ::4], z["cls_L2_strided"], "cls L2")
    for lvl in (0, 1, 2):
        close(out["encoder_features"][lvl + 1][:, ::4, ::16], z[f"features_L{lvl}_strided"], f"features L{lvl}")


def test_backbone_matches_reference_modules_golden():
    replay_reference_backbone_golden(torch_ops, "cpu", rtol=1e-5)


def test_derived_weight_caches_die_with_the_weights():
    """BN-folded / packed weight caches of the SA modules are rebuilt after load_state_dict, an in-place parameter update,
    .to() / .float() and .train() — a stale cache would silently compute with the old weights."""
    from pdanet_b200.config import load_config
    from pdanet_b200.iassd import build_model
    from oracle import torch_ops
    cfg = load_config("kitti")
    torch.manual_seed(0)
    model = build_model(cfg, ops=torch_ops, nms_utils=torch_ops.nms_utils, batched_post_processing=False).eval()
    plain, pda = model.backbone_3d.SA_modules[0], model.backbone_3d.SA_modules[1]

    def poison():
        plain._caches_fresh(), pda._caches_fresh()
        plain._folded, plain._wide, pda._plans = {"stale": 1}, {"stale": 1}, {"stale": 1}

    def clean():
        plain._caches_fresh(), pda._caches_fresh()
        return plain._folded is None and plain._wide == {} and pda._plans == {}

    poison()
    plain._caches_fresh(), pda._caches_fresh()
    assert plain._folded == {"stale": 1} and pda._plans == {"stale": 1}      # nothing changed: caches kept
    model.load_state_dict(model.state_dict())
    assert clean()
    poison()
    with torch.no_grad():
        next(plain.parameters()).mul_(1.0)                                   # in-place update bumps _version
        next(pda.parameters()).add_(0.0)
    assert clean()
    poison()
    model.double()
    assert clean()
    model.float()
    poison()
    model.train()
    assert plain._folded is None and pda._plans == {}
    model.eval()

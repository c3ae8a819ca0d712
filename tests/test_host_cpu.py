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


def replay_reference_backbone_golden(ops_base, device, rtol, name="kitti"):
    """Our backbone vs tests/golden/ref_backbone_<name>.npz — outputs of the REFERENCE's own IASSD_Backbone /
    pointnet2_modules.py run over the oracle ops (tests/golden/make_module_golden.py).  Weights come from the same
    seed (the generator asserts the seeded state_dicts are identical); the class-aware layers replay the reference's
    picks because torch.topk's tie order is unspecified."""
    import numpy as np
    from conftest import GOLDEN
    from pdanet_b200.iassd_backbone import IASSD_Backbone
    z = np.load(GOLDEN / f"ref_backbone_{name}.npz")
    cfg = load_config(name)
    torch.manual_seed(int(z["seed"]))
    bb = IASSD_Backbone(cfg.MODEL.BACKBONE_3D, num_class=len(cfg.CLASS_NAMES), input_channels=4, ops=ops_base).eval().to(device)
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


@pytest.mark.parametrize("name", ["kitti", "once"])
def test_backbone_matches_reference_modules_golden(name):
    """kitti: two 16384-point scenes; once: one 65536-point scene (three wide L5 scales, nsample 64 among them)."""
    replay_reference_backbone_golden(torch_ops, "cpu", rtol=1e-5, name=name)


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


def replay_reference_head_golden(name, ops, nms_utils, device, rtol=1e-5):
    """Head + box decode + post-processing vs tests/golden/ref_head_<name>.npz — outputs of the REFERENCE's own
    `IASSD_Head`, `PointResidual_BinOri_Coder.decode_torch`, `Detector3DTemplate.post_processing` and
    `class_agnostic_nms` (tests/golden/make_head_golden.py imports them unchanged).  The head runs on its seeded inputs;
    every post-processing flavour then runs on the GOLDEN head outputs, so the NMS decisions are made on identical
    numbers and the kept sets must agree exactly."""
    import numpy as np
    from conftest import GOLDEN
    from util import head_golden_inputs, seeded_head_model
    z = np.load(GOLDEN / f"ref_head_{name}.npz")
    cfg = load_config(name)
    model = seeded_head_model(cfg, ops=ops, nms_utils=nms_utils).to(device)
    batch = head_golden_inputs(cfg, name, model.backbone_3d.num_point_features)
    batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}
    B = batch["batch_size"]
    with torch.no_grad():
        out = model.point_head(dict(batch))

    def close(got, want, what):
        got, want = got.detach().cpu(), torch.from_numpy(want)
        err = (got - want).abs().max().item()
        assert err <= rtol * (want.abs().max().item() + 1e-12), f"{what}: {err:.3e}"

    close(out["batch_cls_preds"], z["batch_cls_preds"], "head class logits")
    close(out["batch_box_preds"], z["batch_box_preds"], "decoded boxes")

    feed = dict(out)
    feed["batch_cls_preds"] = torch.from_numpy(z["batch_cls_preds"]).to(device)
    feed["batch_box_preds"] = torch.from_numpy(z["batch_box_preds"]).to(device)
    flavours = [("post_processing", lambda d: model.post_processing(d)[0])]
    if hasattr(model.nms_utils, "nms_batched"):
        flavours.append(("post_processing_batched", lambda d: model.post_processing_batched(d)[0]))
        flavours.append(("post_processing_padded", lambda d: model.unpack_padded(
            {k: v.cpu() for k, v in model.post_processing_padded(d).items()})))
    def table(boxes, scores, labels):
        """(n, 9) rows [box(7), score, label] in a canonical order (scores tie massively: order by the box itself)."""
        t = torch.cat([torch.as_tensor(boxes, dtype=torch.float64).reshape(-1, 7),
                       torch.as_tensor(scores, dtype=torch.float64).reshape(-1, 1),
                       torch.as_tensor(labels, dtype=torch.float64).reshape(-1, 1)], dim=1)
        order = sorted(range(t.shape[0]), key=lambda i: tuple(round(x, 3) for x in t[i, :7].tolist()))
        return t[order]

    for what, fn in flavours:
        with torch.no_grad():
            preds = fn(dict(feed))
        assert len(preds) == B
        for s in range(B):
            got = table(preds[s]["pred_boxes"].cpu(), preds[s]["pred_scores"].cpu(), preds[s]["pred_labels"].cpu())
            want = table(z[f"pred_boxes_{s}"], z[f"pred_scores_{s}"], z[f"pred_labels_{s}"])
            assert got.shape == want.shape, f"{what}, scene {s}: {got.shape[0]} detections kept, the reference keeps {want.shape[0]}"
            # same boxes and labels (inputs are identical), scores to an ulp of the sigmoid (libm vs libdevice)
            assert torch.equal(got[:, 8], want[:, 8]), f"{what}, scene {s}: labels differ"
            assert torch.allclose(got[:, :7], want[:, :7], rtol=0, atol=1e-6), f"{what}, scene {s}: kept boxes differ"
            assert torch.allclose(got[:, 7], want[:, 7], rtol=1e-6, atol=1e-7), f"{what}, scene {s}: scores differ"


@pytest.mark.parametrize("name", ["kitti", "once"])
def test_head_and_post_processing_match_reference_golden(name):
    replay_reference_head_golden(name, torch_ops, torch_ops.nms_utils, "cpu")


def test_custom_ops_registered_with_schemas_and_fake_kernels():
    """`torch.ops.pdab.*` (pdanet_b200/custom_ops.py): schemas registered, shape inference runs without a GPU (FakeTensor),
    and a CPU tensor is refused by the dispatcher (there is no CPU kernel to fall back to)."""
    import pdanet_b200.custom_ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("furthest_point_sample", "furthest_point_sample_with_dist", "ball_query", "gather_operation",
                 "grouping_operation", "nms_keep"):
        assert hasattr(torch.ops.pdab, name)
    with FakeTensorMode():
        xyz = torch.empty(2, 100, 3, device="cuda")
        feats = torch.empty(2, 8, 100, device="cuda")
        idx = torch.ops.pdab.ball_query(0.5, 16, xyz, xyz[:, :7])
        assert idx.shape == (2, 7, 16) and idx.dtype == torch.int32
        assert torch.ops.pdab.furthest_point_sample(xyz, 10).shape == (2, 10)
        assert torch.ops.pdab.grouping_operation(feats, idx).shape == (2, 8, 7, 16)
        keep, num = torch.ops.pdab.nms_keep(torch.empty(50, 7, device="cuda"), 0.1)
        assert keep.shape == (50,) and keep.dtype == torch.int64 and num.shape == (1,)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.pdab.furthest_point_sample(torch.zeros(1, 8, 3), 4)

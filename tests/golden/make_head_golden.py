"""Golden vectors for the LAST third of the path from the REFERENCE's own Python: `IASSD_Head.forward`
(pcdet/models/dense_heads/IASSD_head.py:1343-1399), `PointResidual_BinOri_Coder.decode_torch`
(pcdet/utils/box_coder_utils.py:279-319), `Detector3DTemplate.post_processing`
(pcdet/models/detectors/detector3d_template.py:179-285) and `class_agnostic_nms`
(pcdet/models/model_utils/model_nms_utils.py:6-25) -> `iou3d_nms_utils.nms_gpu` (IOU/iou3d_nms_utils.py:84-99).

Runs only in the build container (needs /root/reference).  The reference files are imported UNCHANGED through the
stub loader of make_module_golden.py; packages they import but never touch on this path (spconv, SharedArray,
roiaware_pool3d_cuda, the other detectors' backbones) are empty stubs, `iou3d_nms_cuda` is the CPU oracle (same pybind
names), and `Tensor.cuda()` is the identity while the reference code runs (the coder's constructor and `nms_gpu` call it).
The head of the reference is built from the REFERENCE's yaml (`tools/cfgs/{kitti,once}_models/PDA-SSD.yaml`) under
torch.manual_seed(0); ours from our cfg under the same seed — the state_dicts must be identical.

Stored in tests/golden/ref_head_{kitti,once}.npz: `batch_cls_preds`, `batch_box_preds` and per scene the kept boxes /
scores / labels of the reference's post-processing, for the seeded head and inputs of tests/util.py
(`seeded_head_model`, `head_golden_inputs`).  tests/test_host_cpu.py replays them on our mirror with the CPU oracle ops, tests/test_gpu_model.py on the
CUDA path (`post_processing`, `_batched`, `_padded`).

    python tests/golden/make_head_golden.py [kitti|once]
"""
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from make_module_golden import _pkg, import_reference  # noqa: E402
from pdanet_b200.config import AttrDict, load_config  # noqa: E402
from pdanet_b200.iassd import build_model  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference_head():
    import_reference()                                   # pcdet, pcdet.ops.pointnet2..., pcdet.models, backbones_3d
    _stub("SharedArray")
    _stub("torch_scatter", scatter_mean=None, scatter_max=None)   # cluster_contrastloss.py:6 (a training loss)
    _pkg("pcdet.utils", REF / "pcdet/utils")
    _stub("pcdet.utils.spconv_utils", find_all_spconv_keys=lambda *a, **k: set())
    _pkg("pcdet.ops.roiaware_pool3d", REF / "pcdet/ops/roiaware_pool3d")
    _stub("pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda")
    _pkg("pcdet.ops.iou3d_nms", REF / "pcdet/ops/iou3d_nms")
    sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_cuda"] = oracle   # nms_gpu(boxes, keep, thresh) etc. on CPU tensors
    _pkg("pcdet.models.dense_heads", REF / "pcdet/models/dense_heads")
    _pkg("pcdet.models.model_utils", REF / "pcdet/models/model_utils")
    _pkg("pcdet.models.detectors", REF / "pcdet/models/detectors")
    # detector3d_template.py:8-10 imports the sibling packages of every other detector; none is used by post_processing
    bb3 = sys.modules["pcdet.models.backbones_3d"]
    bb3.pfe = _stub("pcdet.models.backbones_3d.pfe")
    bb3.vfe = _stub("pcdet.models.backbones_3d.vfe")
    bb2 = _stub("pcdet.models.backbones_2d")
    bb2.map_to_bev = _stub("pcdet.models.backbones_2d.map_to_bev")
    _stub("pcdet.models.roi_heads")
    models = sys.modules["pcdet.models"]
    models.backbones_2d, models.backbones_3d, models.roi_heads = bb2, bb3, sys.modules["pcdet.models.roi_heads"]
    models.dense_heads = sys.modules["pcdet.models.dense_heads"]
    head = importlib.import_module("pcdet.models.dense_heads.IASSD_head")
    det = importlib.import_module("pcdet.models.detectors.detector3d_template")
    return head, det


class cuda_is_identity:
    """The reference calls `.cuda()` on the coder's mean sizes (box_coder_utils.py:233) and on the NMS keep list
    (iou3d_nms_utils.py:99); on the CPU oracle both are no-ops."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig


def main(name="kitti"):
    head_mod, det_mod = import_reference_head()
    ref_yaml = REF / "tools/cfgs" / f"{name}_models" / "PDA-SSD.yaml"
    ref_cfg = AttrDict(yaml.safe_load(open(ref_yaml)))
    cfg = load_config(name)
    num_class = len(cfg.CLASS_NAMES)
    assert list(ref_cfg.CLASS_NAMES) == list(cfg.CLASS_NAMES)

    from util import head_golden_inputs, seeded_head_model
    ours = seeded_head_model(cfg)
    in_dim = ours.backbone_3d.num_point_features
    with cuda_is_identity():
        torch.manual_seed(1)
        ref_head = head_mod.IASSD_Head(num_class=num_class, input_channels=in_dim, model_cfg=ref_cfg.MODEL.POINT_HEAD).eval()
    # our head, built under the same seed, must be the reference's head parameter for parameter
    sd_ref, sd_ours = ref_head.state_dict(), ours.point_head.state_dict()
    extra = [k for k in sd_ref if k not in sd_ours and not k.startswith(("cls_loss_func", "reg_loss_func", "nce"))]
    assert not extra, f"reference head parameters without a counterpart: {extra}"
    for k, v in sd_ours.items():
        if not (k.endswith("running_mean") or k.endswith("running_var")):
            assert torch.equal(sd_ref[k], v), f"seeded init differs at {k}"
    ref_head.load_state_dict(sd_ours, strict=False)      # (the randomised BatchNorm statistics)
    print(f"head state_dict identical: {len(sd_ours)} tensors")
    batch = head_golden_inputs(cfg, name, in_dim)
    B = batch["batch_size"]

    with torch.no_grad(), cuda_is_identity():
        out = ref_head(dict(batch))
        # a detector object is only needed for its two attributes; post_processing is called unbound
        shell = types.SimpleNamespace(model_cfg=ref_cfg.MODEL, num_class=num_class,
                                      generate_recall_record=det_mod.Detector3DTemplate.generate_recall_record)
        preds, _ = det_mod.Detector3DTemplate.post_processing(shell, out)
        mine = ours.point_head(dict(batch))
        my_preds, _ = ours.post_processing(mine)
    assert torch.allclose(out["batch_cls_preds"], mine["batch_cls_preds"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(out["batch_box_preds"], mine["batch_box_preds"], rtol=1e-5, atol=1e-5)
    for s in range(B):
        for k in ("pred_boxes", "pred_scores", "pred_labels"):
            assert preds[s][k].shape == my_preds[s][k].shape, (s, k, preds[s][k].shape, my_preds[s][k].shape)
            assert torch.allclose(preds[s][k].float(), my_preds[s][k].float(), rtol=1e-5, atol=1e-5), (s, k)
    print(f"{name}: our head + post-processing reproduce the reference's on this input; kept per scene:",
          [int(p["pred_boxes"].shape[0]) for p in preds])

    fx = {"batch": np.int64(B), "batch_cls_preds": out["batch_cls_preds"].numpy(),
          "batch_box_preds": out["batch_box_preds"].numpy()}
    for s in range(B):
        fx[f"pred_boxes_{s}"] = preds[s]["pred_boxes"].numpy()
        fx[f"pred_scores_{s}"] = preds[s]["pred_scores"].numpy()
        fx[f"pred_labels_{s}"] = preds[s]["pred_labels"].numpy().astype(np.int64)
    path = ROOT / "tests/golden" / f"ref_head_{name}.npz"
    np.savez_compressed(path, **fx)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "kitti")

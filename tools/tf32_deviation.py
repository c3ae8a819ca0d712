#!/usr/bin/env python
"""How far do the unchanged PyTorch layers move when fp32 matmuls run as TF32?  Backbone features and decoded boxes,
TF32 vs IEEE fp32 on the same GPU, teacher-forced on identical sampled points (so no top-k flips)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pdanet_b200.config import load_config  # noqa: E402
from pdanet_b200.iassd import build_model  # noqa: E402
from pdanet_b200.synthetic import make_batch  # noqa: E402

cfg = load_config("kitti")
torch.manual_seed(0)
model = build_model(cfg).cuda().eval()
pts = make_batch(4, 16384, cfg.POINT_CLOUD_RANGE)["points"].cuda()


def run(tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        d = model.backbone_3d({"batch_size": 4, "points": pts})
        d = model.point_head(d)
    return d


a, b = run(False), run(True)
for lvl in (1, 2, 3):
    fa, fb = a["encoder_features"][lvl], b["encoder_features"][lvl]
    same = torch.equal(a["encoder_xyz"][lvl], b["encoder_xyz"][lvl])
    err = (fa - fb).abs().max().item() / fa.abs().max().item() if same else float("nan")
    print(f"L{lvl - 1} features: same points {same}, max rel-to-max err {err:.2e}")
same = torch.equal(a["centers"], b["centers"])
print("centers identical:", same)
if same:
    for k in ("centers_features", "batch_cls_preds", "batch_box_preds"):
        print(k, f"{((a[k] - b[k]).abs().max() / a[k].abs().max()).item():.2e}")

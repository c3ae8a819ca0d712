#!/usr/bin/env python
"""Two sequential, un-pipelined inference steps (KITTI config, 16 x 16384 points) launched eagerly, so every kernel shows
up as its own launch: target for `ncu -k regex:... -s <launches of step 1> -c <launches of step 2>`.
    python tools/run_step.py [--tc-passes 2]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from pdanet_b200 import _lib  # noqa: E402
from pdanet_b200.config import load_config  # noqa: E402
from pdanet_b200.runner import SceneRunner  # noqa: E402
from pdanet_b200.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tc-passes", type=int, default=None)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
cfg = load_config("kitti")
runner = SceneRunner(cfg, device="cuda:0", batch_size=16, num_points=16384, seed=0, tc_passes=args.tc_passes)
pts = make_batch(16, 16384, cfg.POINT_CLOUD_RANGE)["points"].cuda()
for _ in range(args.steps):
    runner.infer_device(pts)
    torch.cuda.synchronize()
print("launches through the C ABI:", dict(_lib.launch_counts))

#!/usr/bin/env python
"""Per-kernel device time of one bench step via torch.profiler (CUPTI) — cheap iteration aid; the committed
launch lists under profiles/ come from ncu.    python tools/profile_step.py [kitti|once] [batch] [tc_passes]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pdanet_b200.runner import SceneRunner  # noqa: E402
from pdanet_b200.synthetic import make_batch  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "kitti"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
tc_passes = int(sys.argv[3]) if len(sys.argv) > 3 else None
runner = SceneRunner(cfg_name, batch_size=batch, tc_passes=tc_passes)
pts = make_batch(batch, runner.num_points, runner.cfg.POINT_CLOUD_RANGE)["points"].cuda()
for _ in range(3):
    runner.infer_device(pts)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner.infer_device(pts)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0
        and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print(f"device time of one step: {total:.2f} ms over {sum(r[2] for r in rows)} kernels")
for name, ms, n in rows[:70]:
    print(f"{ms:8.3f} ms {100 * ms / total:5.1f}% n={n:4d}  {name[:110]}")

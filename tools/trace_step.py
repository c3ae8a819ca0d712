#!/usr/bin/env python
"""Ordered kernel list of ONE sequential step (torch.profiler / CUPTI): which launches are ours, which are torch glue, and
where the glue sits.    python tools/trace_step.py [kitti|once] [batch] > gpurun_out/trace.txt"""
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
runner = SceneRunner(cfg_name, batch_size=batch)
pts = make_batch(batch, runner.num_points, runner.cfg.POINT_CLOUD_RANGE)["points"].cuda()
for _ in range(3):
    runner.infer_device(pts)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=False, record_shapes=True) as prof:
    with torch.profiler.record_function("STEP"):
        runner.infer_device(pts)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  {e.device_time_total:8.1f} us  {e.name[:120]}")

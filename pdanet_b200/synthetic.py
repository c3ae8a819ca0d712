"""Synthetic point clouds in the reference's batch layout.

`points` is (B*N, 5) = [batch_idx, x, y, z, intensity], scenes contiguous, as produced by the reference's
`DatasetTemplate.collate_batch` (pcdet/datasets/dataset.py:173-178).  Coordinates are uniform in the
dataset's POINT_CLOUD_RANGE, intensity uniform in [0,1); `duplicate_frac` re-uses earlier points of a
scene to mimic the padding-by-duplication of short scenes (data_processor.py:212-214), which creates
exact distance ties."""
from __future__ import annotations

import torch


def make_scene(scene_id: int, n: int, pc_range, duplicate_frac: float = 0.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + scene_id)
    lo = torch.tensor(pc_range[:3], dtype=torch.float32)
    hi = torch.tensor(pc_range[3:], dtype=torch.float32)
    xyz = lo + (hi - lo) * torch.rand(n, 3, generator=g)
    inten = torch.rand(n, 1, generator=g)
    pts = torch.cat([xyz, inten], dim=1)
    ndup = int(n * duplicate_frac)
    if ndup > 0:
        src = torch.randint(0, n - ndup, (ndup,), generator=g)
        pts[n - ndup:] = pts[src]
    return pts


def make_batch(batch_size: int, n: int, pc_range, first_scene: int = 0, duplicate_frac: float = 0.0) -> dict:
    scenes = [make_scene(first_scene + s, n, pc_range, duplicate_frac) for s in range(batch_size)]
    bidx = torch.arange(batch_size, dtype=torch.float32).repeat_interleave(n).unsqueeze(1)
    return {"batch_size": batch_size, "points": torch.cat([bidx, torch.cat(scenes, dim=0)], dim=1).contiguous()}

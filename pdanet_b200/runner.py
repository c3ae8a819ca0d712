"""Scene-batch inference runner: the call a user makes to push host point clouds through PDA-SSD.

`SceneRunner.infer(points_host)` takes the reference's batch layout — a (B*N, 5) fp32 array
[batch_idx, x, y, z, intensity] in host memory (pcdet/datasets/dataset.py:173-178) — stages it through a
pinned buffer, runs the detector on the runner's device / stream and returns the per-scene predictions
as host tensors.  Scenes are independent: with several GPUs each rank owns a
`SceneRunner` and a disjoint slice of the scenes (`shard_scenes`), with no collective on the data path.
"""
from __future__ import annotations

from typing import List

import torch

from .config import load_config
from .iassd import build_model


def shard_scenes(num_scenes: int, world_size: int, rank: int) -> range:
    """Contiguous block of scene ids owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(num_scenes, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class SceneRunner:
    def __init__(self, cfg="kitti", device="cuda:0", batch_size=16, num_points=None, seed=0, model=None):
        self.cfg = load_config(cfg) if isinstance(cfg, str) else cfg
        self.device = torch.device(device)
        self.batch_size = batch_size
        self.num_points = num_points or self.cfg.NUM_POINTS
        if model is None:
            torch.manual_seed(seed)
            model = build_model(self.cfg)
        self.model = model.to(self.device).eval()
        rows = self.batch_size * self.num_points
        self._pinned = torch.empty(rows, 5, dtype=torch.float32).pin_memory()
        self._dev = torch.empty(rows, 5, dtype=torch.float32, device=self.device)
        self.h2d_bytes = self._pinned.numel() * 4
        self.d2h_bytes = 0

    @torch.no_grad()
    def infer_device(self, points_dev: torch.Tensor, batch_size=None):
        """points already in HBM -> pred_dicts on the device."""
        return self.model({"batch_size": batch_size or self.batch_size, "points": points_dev})[0]

    @torch.no_grad()
    def infer(self, points_host: torch.Tensor) -> List[dict]:
        """Host (B*N,5) -> host predictions; H2D and D2H copies included."""
        assert points_host.shape == self._pinned.shape and points_host.dtype == torch.float32
        if points_host.is_pinned():
            src = points_host
        else:
            self._pinned.copy_(points_host)
            src = self._pinned
        with torch.cuda.device(self.device):
            self._dev.copy_(src, non_blocking=True)
            preds = self.infer_device(self._dev)
            out, nbytes = [], 0
            for p in preds:
                host = {k: v.to("cpu", non_blocking=True) for k, v in p.items()}
                nbytes += sum(v.numel() * v.element_size() for v in host.values())
                out.append(host)
            torch.cuda.current_stream().synchronize()
        self.d2h_bytes = nbytes
        return out

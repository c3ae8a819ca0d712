"""Host-side mirror of the part of the reference's `roiaware_pool3d_utils` that PDA-SSD uses
(pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:28-41): `points_in_boxes_gpu`, the native op behind the detection
head's training-time target assignment (pcdet/models/dense_heads/IASSD_head.py:169,196,214).  CUDA tensors only."""
from __future__ import annotations

import torch

from . import _lib


def points_in_boxes_gpu(points: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """points (B,M,3), boxes (B,T,7) -> box_idxs_of_pts (B,M) int32, -1 = background, else the first containing box."""
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    if not (points.is_cuda and boxes.is_cuda):
        raise RuntimeError("points_in_boxes_gpu needs CUDA tensors (pdanet_b200 has no CPU path)")
    batch_size, num_points, _ = points.shape
    box_idxs_of_pts = torch.full((batch_size, num_points), -1, dtype=torch.int32, device=points.device)
    boxes, points = boxes.contiguous().float(), points.contiguous().float()
    with torch.cuda.device(points.device):
        _lib.call("pdab_points_in_boxes", batch_size, boxes.shape[1], num_points, boxes.data_ptr(), points.data_ptr(),
                  box_idxs_of_pts.data_ptr(), torch.cuda.current_stream(points.device).cuda_stream)
    return box_idxs_of_pts

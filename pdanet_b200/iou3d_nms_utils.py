"""Host-side mirror of the reference's `iou3d_nms_utils` (IOU/iou3d_nms_utils.py) over libpdab.so.

`nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs) -> (LongTensor idx, None)` keeps the
reference's signature and return convention (IOU/iou3d_nms_utils.py:84-99) because the caller
invokes it as `getattr(iou3d_nms_utils, cfg.NMS_TYPE)(boxes, scores, thresh, **cfg)`
(model_utils/model_nms_utils.py:17-19).  `nms_batched` is the additive on-device form used by the
detector's batched post-processing: one launch pair for all scenes, no host round trip.
"""
from __future__ import annotations

import torch

from . import _lib
from . import iou3d_nms_cuda


def boxes_iou_bev(boxes_a, boxes_b):
    """(N,7) x (M,7) -> (N,M) rotated BEV IoU.  IOU/iou3d_nms_utils.py:31-45."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans = torch.zeros(boxes_a.shape[0], boxes_b.shape[0], dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7) x (M,7) -> (N,M) 3D IoU = BEV overlap x height overlap / union.  IOU/iou3d_nms_utils.py:48-81."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_max = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    a_min = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_max = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    b_min = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = torch.zeros(boxes_a.shape[0], boxes_b.shape[0], dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), overlaps_bev)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """Rotated NMS.  boxes (N,7), scores (N) -> (indices into the input kept, None)."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.empty(boxes.size(0), dtype=torch.int64)
    num_out = iou3d_nms_cuda.nms_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """Axis-aligned variant.  IOU/iou3d_nms_utils.py:102-116."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    boxes = boxes[order].contiguous()
    keep = torch.empty(boxes.size(0), dtype=torch.int64)
    num_out = iou3d_nms_cuda.nms_normal_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_batched(boxes: torch.Tensor, counts: torch.Tensor, thresh: float):
    """boxes (S, stride, 7) fp32 cuda, each scene's first counts[s] rows valid and sorted by score desc;
    returns (keep (S, stride) int64 positions, num_keep (S) int32), both on the device, no sync."""
    assert boxes.is_cuda and boxes.is_contiguous() and boxes.dtype == torch.float32 and boxes.shape[-1] == 7
    S, stride, _ = boxes.shape
    counts = counts.to(device=boxes.device, dtype=torch.int32).contiguous()
    keep = torch.empty(S, stride, dtype=torch.int64, device=boxes.device)
    num = torch.empty(S, dtype=torch.int32, device=boxes.device)
    ws_bytes = _lib.lib().pdab_nms_workspace_bytes(stride) * S
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _lib.call("pdab_nms_batched", boxes.data_ptr(), counts.data_ptr(), S, stride, float(thresh), keep.data_ptr(),
                  num.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(boxes.device).cuda_stream)
    return keep, num

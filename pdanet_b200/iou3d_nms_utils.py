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


def post_front(cls: torch.Tensor, boxes: torch.Tensor, batch_size: int, normalized: bool, score_thresh: float, pre_max: int):
    """One kernel for the front half of the batched post-processing (`pdab_post_front`): cls (B*M, num_class) logits (row
    stride free), boxes (B*M, >= 7) -> order (B, M) int32, sorted_boxes (B, M, 7), counts (B) int32, scores (B, M),
    raw_max (B, M), labels (B, M) int64."""
    assert cls.is_cuda and boxes.is_cuda and cls.dtype == boxes.dtype == torch.float32
    assert cls.stride(1) == 1 and boxes.stride(1) == 1
    M = cls.shape[0] // batch_size
    dev = cls.device
    order = torch.empty(batch_size, M, dtype=torch.int32, device=dev)
    sorted_boxes = torch.empty(batch_size, M, 7, dtype=torch.float32, device=dev)
    counts = torch.empty(batch_size, dtype=torch.int32, device=dev)
    scores = torch.empty(batch_size, M, dtype=torch.float32, device=dev)
    raw_max = torch.empty(batch_size, M, dtype=torch.float32, device=dev)
    labels = torch.empty(batch_size, M, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("pdab_post_front", batch_size, M, cls.shape[1], cls.stride(0), boxes.stride(0), int(bool(normalized)),
                  float(score_thresh), int(pre_max), cls.data_ptr(), boxes.data_ptr(), order.data_ptr(),
                  sorted_boxes.data_ptr(), counts.data_ptr(), scores.data_ptr(), raw_max.data_ptr(), labels.data_ptr(),
                  torch.cuda.current_stream(dev).cuda_stream)
    return order, sorted_boxes, counts, scores, raw_max, labels


def post_select(keep, num_keep, order, boxes, scores, labels, P: int):
    """Back half (`pdab_post_select`): padded (B, P, nb) boxes, (B, P) scores / labels and num (B), zeros behind num."""
    B, M = order.shape
    nb = boxes.shape[1]
    dev = boxes.device
    out_boxes = torch.empty(B, P, nb, dtype=torch.float32, device=dev)
    out_scores = torch.empty(B, P, dtype=torch.float32, device=dev)
    out_labels = torch.empty(B, P, dtype=torch.int64, device=dev)
    out_num = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("pdab_post_select", B, M, P, nb, boxes.stride(0), keep.data_ptr(), num_keep.data_ptr(), order.data_ptr(),
                  boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), out_boxes.data_ptr(), out_scores.data_ptr(),
                  out_labels.data_ptr(), out_num.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return out_boxes, out_scores, out_labels, out_num

"""Drop-in for the reference pybind11 module `iou3d_nms_cuda`
(IOU/src/iou3d_nms_api.cpp:11-17, IOU = pcdet/ops/iou3d_nms in the reference).

Same names, arities and the caller-allocates convention, including nms_gpu's
CPU int64 `keep` tensor and integer return (IOU/src/iou3d_nms.cpp:90-136).
Bad inputs raise RuntimeError instead of exit(-1) (IOU/src/iou3d_nms.cpp:14-26).
boxes_iou_bev_cpu is not provided: the product has no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from .pointnet2_batch_cuda import _chk, _same_device, _stream


def _boxes(t, name):
    if t.dim() != 2 or t.shape[1] != 7:
        raise RuntimeError(f"{name} must have shape (N, 7)")
    return _chk(t, name, torch.float32)


def _nms(boxes, keep, thresh, normal):
    pb = _boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous() or keep.numel() < boxes.shape[0]:
        raise RuntimeError("keep must be a contiguous CPU int64 tensor with at least N elements")
    n = boxes.shape[0]
    with torch.cuda.device(boxes.device):
        rc = _lib.lib().pdab_nms_host(pb, n, float(thresh), keep.data_ptr(), int(normal), _stream(boxes))
    if rc < 0:
        raise _lib.PdabError("pdab_nms_host", rc, _lib.lib().pdab_error_string(rc).decode())
    return rc


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """boxes (N,7) cuda fp32 sorted by score desc; keep (N) CPU int64; returns num_to_keep."""
    return _nms(boxes, keep, nms_overlap_thresh, False)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, True)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    pa, pb = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    po = _chk(ans_overlap, "ans_overlap", torch.float32, (boxes_a.shape[0], boxes_b.shape[0]))
    with _same_device(boxes_a, boxes_b, ans_overlap):
        _lib.call("pdab_boxes_overlap_bev", boxes_a.shape[0], pa, boxes_b.shape[0], pb, po, _stream(boxes_a))
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    pa, pb = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    po = _chk(ans_iou, "ans_iou", torch.float32, (boxes_a.shape[0], boxes_b.shape[0]))
    with _same_device(boxes_a, boxes_b, ans_iou):
        _lib.call("pdab_boxes_iou_bev", boxes_a.shape[0], pa, boxes_b.shape[0], pb, po, _stream(boxes_a))
    return 1


def boxes_iou_bev_cpu(*_a, **_k):
    raise NotImplementedError("pdanet_b200 has no CPU path; the CPU IoU lives in oracle/ (test infrastructure)")

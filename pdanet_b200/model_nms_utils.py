"""Mirror of the reference's class-agnostic NMS front end (pcdet/models/model_utils/model_nms_utils.py:6-25)."""
from __future__ import annotations

import torch

from . import iou3d_nms_utils as _cuda_nms


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None, nms_utils=None):
    """score mask -> topk(NMS_PRE_MAXSIZE) -> NMS -> first NMS_POST_MAXSIZE.  Returns
    (indices into the unmasked input, their scores)."""
    nms_utils = nms_utils if nms_utils is not None else _cuda_nms
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = box_scores >= score_thresh
        box_scores = box_scores[scores_mask]
        box_preds = box_preds[scores_mask]
    selected = []
    if box_scores.shape[0] > 0:
        top_scores, indices = torch.topk(box_scores, k=min(nms_config.NMS_PRE_MAXSIZE, box_scores.shape[0]))
        keep_idx, _ = getattr(nms_utils, nms_config.NMS_TYPE)(
            box_preds[indices][:, 0:7], top_scores, nms_config.NMS_THRESH, **nms_config)
        selected = indices[keep_idx[:nms_config.NMS_POST_MAXSIZE]]
    if score_thresh is not None:
        original = scores_mask.nonzero().view(-1)
        selected = original[selected]
    return selected, src_box_scores[selected]

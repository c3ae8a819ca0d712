"""PDA-SSD detector (MODEL.NAME: IASSD): backbone -> head -> post-processing.

Mirror of the reference's `IASSD` / `Detector3DTemplate` inference path
(pcdet/models/detectors/IASSD.py:8-20, detector3d_template.py:179-285): module attribute names
`backbone_3d` and `point_head` and the `global_step` buffer match, so `state_dict()` keys are the
reference's.  `post_processing` keeps the reference's per-scene semantics and return value
(`pred_dicts`, `recall_dict`); `post_processing_batched` produces the same `pred_dicts` with one
batched sort + one batched on-device NMS and a single host sync for the whole batch.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import iou3d_nms_utils as _cuda_nms
from . import model_nms_utils
from .iassd_backbone import IASSD_Backbone
from .iassd_head import IASSD_Head


class IASSD(nn.Module):
    def __init__(self, model_cfg, num_class, num_point_features=4, ops=None, nms_utils=None,
                 batched_post_processing=True):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_class = num_class
        self.register_buffer("global_step", torch.zeros(1, dtype=torch.long))
        self.backbone_3d = IASSD_Backbone(model_cfg.BACKBONE_3D, num_class=num_class,
                                          input_channels=num_point_features, ops=ops)
        self.point_head = IASSD_Head(num_class=num_class, input_channels=self.backbone_3d.num_point_features,
                                     model_cfg=model_cfg.POINT_HEAD)
        self.module_list = [self.backbone_3d, self.point_head]
        self.nms_utils = nms_utils if nms_utils is not None else _cuda_nms
        self.batched_post_processing = batched_post_processing and hasattr(self.nms_utils, "nms_batched")
        self.output_padded = False  # True: forward returns post_processing_padded's fixed-shape device tensors
        self.fused_post_processing = True   # post_processing_padded: two kernels around the batched NMS (csrc/post.cu)

    def forward(self, batch_dict):
        for module in self.module_list:
            batch_dict = module(batch_dict)
        if self.training:
            # Train mode: every op of the path differentiates (gather / group / interpolate through the C ABI's gradient entry
            # points, deterministic by default; the PDA block in the reference's statement order on autograd) and the head's
            # raw predictions are returned for the caller's objective.  The reference's target assignment and loss terms
            # (pcdet/models/dense_heads/IASSD_head.py:169-1330) are NOT ported: `point_head.get_loss` raises.
            return batch_dict
        if self.output_padded:
            return self.post_processing_padded(batch_dict)
        if self.batched_post_processing:
            return self.post_processing_batched(batch_dict)
        return self.post_processing(batch_dict)

    # ------------------------------------------------------------------ reference semantics
    def post_processing(self, batch_dict):
        cfg = self.model_cfg.POST_PROCESSING
        batch_size = batch_dict["batch_size"]
        pred_dicts = []
        for index in range(batch_size):
            mask = batch_dict["batch_index"] == index
            box_preds = batch_dict["batch_box_preds"][mask]
            src_cls = batch_dict["batch_cls_preds"][mask]
            assert src_cls.shape[1] in [1, self.num_class]
            cls_preds = src_cls if batch_dict["cls_preds_normalized"] else torch.sigmoid(src_cls)
            if cfg.NMS_CONFIG.MULTI_CLASSES_NMS:
                raise NotImplementedError("PDA-SSD uses class-agnostic NMS")
            cls_preds, label_preds = torch.max(cls_preds, dim=-1)
            label_preds = label_preds + 1
            selected, selected_scores = model_nms_utils.class_agnostic_nms(
                box_scores=cls_preds, box_preds=box_preds, nms_config=cfg.NMS_CONFIG,
                score_thresh=cfg.SCORE_THRESH, nms_utils=self.nms_utils)
            if cfg.OUTPUT_RAW_SCORE:
                selected_scores = torch.max(src_cls, dim=-1)[0][selected]
            pred_dicts.append({"pred_boxes": box_preds[selected], "pred_scores": selected_scores,
                               "pred_labels": label_preds[selected]})
        return pred_dicts, {}

    # ------------------------------------------------------------------ batched, device-side
    def post_processing_batched(self, batch_dict):
        cfg = self.model_cfg.POST_PROCESSING
        nms_cfg = cfg.NMS_CONFIG
        B = batch_dict["batch_size"]
        boxes = batch_dict["batch_box_preds"]
        M = boxes.shape[0] // B
        boxes = boxes.view(B, M, boxes.shape[-1])
        src_cls = batch_dict["batch_cls_preds"].view(B, M, -1)
        probs = src_cls if batch_dict["cls_preds_normalized"] else torch.sigmoid(src_cls)
        scores, labels = probs.max(dim=-1)
        labels = labels + 1
        valid = scores >= cfg.SCORE_THRESH
        key = torch.where(valid, scores, torch.full_like(scores, float("-inf")))
        _, order = key.sort(dim=1, descending=True, stable=True)      # (score desc, index asc)
        counts = valid.sum(dim=1).clamp(max=nms_cfg.NMS_PRE_MAXSIZE).int()
        sorted_boxes = torch.gather(boxes[..., :7], 1, order.unsqueeze(-1).expand(-1, -1, 7)).contiguous()
        keep, num = self.nms_utils.nms_batched(sorted_boxes, counts, nms_cfg.NMS_THRESH)
        num_host = num.clamp(max=nms_cfg.NMS_POST_MAXSIZE).cpu().tolist()  # the batch's only host sync
        raw_max = src_cls.max(dim=-1)[0] if cfg.OUTPUT_RAW_SCORE else None
        pred_dicts = []
        for s in range(B):
            sel = order[s].index_select(0, keep[s, :num_host[s]])
            pred_dicts.append({"pred_boxes": boxes[s].index_select(0, sel),
                               "pred_scores": (raw_max if raw_max is not None else scores)[s].index_select(0, sel),
                               "pred_labels": labels[s].index_select(0, sel)})
        return pred_dicts, {}


    # ------------------------------------------------------------------ fixed shapes, no host sync
    def post_processing_padded(self, batch_dict):
        """Same selection as `post_processing_batched`, returned as fixed-shape device tensors with NO host sync, so the
        whole forward can be captured in a CUDA graph and pipelined across streams:
        pred_boxes (B,P,C), pred_scores (B,P), pred_labels (B,P) int64, num (B) int32; rows >= num[s] are zero.
        `unpack_padded` turns the host copy into the reference's list of pred_dicts."""
        cfg = self.model_cfg.POST_PROCESSING
        nms_cfg = cfg.NMS_CONFIG
        B = batch_dict["batch_size"]
        boxes = batch_dict["batch_box_preds"]
        M = boxes.shape[0] // B
        if (self.fused_post_processing and hasattr(self.nms_utils, "post_front") and boxes.is_cuda and M <= 4096
                and boxes.stride(1) == 1 and batch_dict["batch_cls_preds"].stride(1) == 1):
            # the same steps as below in two kernels around the batched NMS (csrc/post.cu), bit-identical results
            order, sorted_boxes, counts, scores, raw_max, labels = self.nms_utils.post_front(
                batch_dict["batch_cls_preds"], boxes, B, batch_dict["cls_preds_normalized"], cfg.SCORE_THRESH,
                nms_cfg.NMS_PRE_MAXSIZE)
            keep, num = self.nms_utils.nms_batched(sorted_boxes, counts, nms_cfg.NMS_THRESH)
            out_boxes, out_scores, out_labels, out_num = self.nms_utils.post_select(
                keep, num, order, boxes, raw_max if cfg.OUTPUT_RAW_SCORE else scores, labels, min(M, nms_cfg.NMS_POST_MAXSIZE))
            return {"pred_boxes": out_boxes, "pred_scores": out_scores, "pred_labels": out_labels, "num": out_num}
        boxes = boxes.view(B, M, boxes.shape[-1])
        src_cls = batch_dict["batch_cls_preds"].view(B, M, -1)
        probs = src_cls if batch_dict["cls_preds_normalized"] else torch.sigmoid(src_cls)
        scores, labels = probs.max(dim=-1)
        labels = labels + 1
        valid = scores >= cfg.SCORE_THRESH
        key = torch.where(valid, scores, torch.full_like(scores, float("-inf")))
        _, order = key.sort(dim=1, descending=True, stable=True)      # (score desc, index asc)
        counts = valid.sum(dim=1).clamp(max=nms_cfg.NMS_PRE_MAXSIZE).int()
        sorted_boxes = torch.gather(boxes[..., :7], 1, order.unsqueeze(-1).expand(-1, -1, 7)).contiguous()
        keep, num = self.nms_utils.nms_batched(sorted_boxes, counts, nms_cfg.NMS_THRESH)
        P = min(M, nms_cfg.NMS_POST_MAXSIZE)
        num = num.clamp(max=P)
        live = torch.arange(P, device=boxes.device).unsqueeze(0) < num.unsqueeze(1)            # (B, P)
        sel = torch.gather(order, 1, torch.where(live, keep[:, :P], torch.zeros_like(keep[:, :P])))
        out_scores = src_cls.max(dim=-1)[0] if cfg.OUTPUT_RAW_SCORE else scores
        # dead slots gather row order[:, 0]; they are SELECTED to zero, not multiplied (0 * inf from an overflowing decode
        # would put NaN into rows the contract promises to be zero)
        g_boxes = torch.gather(boxes, 1, sel.unsqueeze(-1).expand(-1, -1, boxes.shape[-1]))
        g_scores, g_labels = torch.gather(out_scores, 1, sel), torch.gather(labels, 1, sel)
        return {
            "pred_boxes": torch.where(live.unsqueeze(-1), g_boxes, torch.zeros_like(g_boxes)),
            "pred_scores": torch.where(live, g_scores, torch.zeros_like(g_scores)),
            "pred_labels": torch.where(live, g_labels, torch.zeros_like(g_labels)),
            "num": num,
        }

    @staticmethod
    def unpack_padded(padded_host):
        """Host copy of `post_processing_padded`'s tensors -> list of pred_dicts (views, no copies)."""
        nums = padded_host["num"].tolist()
        return [{"pred_boxes": padded_host["pred_boxes"][s, :n], "pred_scores": padded_host["pred_scores"][s, :n],
                 "pred_labels": padded_host["pred_labels"][s, :n]} for s, n in enumerate(nums)]


def build_model(cfg, ops=None, nms_utils=None, batched_post_processing=True) -> IASSD:
    """cfg: pdanet_b200.config.load_config('kitti' | 'once' | path)."""
    return IASSD(cfg.MODEL, num_class=len(cfg.CLASS_NAMES), num_point_features=cfg.get("NUM_POINT_FEATURES", 4),
                 ops=ops, nms_utils=nms_utils, batched_post_processing=batched_post_processing)

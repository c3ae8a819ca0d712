"""PDA-SSD detection head, inference side.

Mirror of `IASSD_Head.forward` / `generate_predicted_boxes` / `make_fc_layers`
(pcdet/models/dense_heads/IASSD_head.py:1343-1399, point_head_template.py:36-47,193-207) with the
reference's sub-module names (`cls_center_layers`, `box_center_layers`) so checkpoints load.
Target assignment and the loss functions are training-only and out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import box_coder_utils


def make_fc_layers(fc_cfg, input_channels, output_channels):
    layers, c_in = [], input_channels
    for width in fc_cfg:
        layers += [nn.Linear(c_in, width, bias=False), nn.BatchNorm1d(width), nn.ReLU()]
        c_in = width
    layers.append(nn.Linear(c_in, output_channels, bias=True))
    return nn.Sequential(*layers)


class IASSD_Head(nn.Module):
    def __init__(self, num_class, input_channels, model_cfg, predict_boxes_when_training=False, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_class = num_class
        self.predict_boxes_when_training = predict_boxes_when_training
        target_cfg = self.model_cfg.TARGET_CONFIG
        self.box_coder = getattr(box_coder_utils, target_cfg.BOX_CODER)(**target_cfg.BOX_CODER_CONFIG)
        dim = self.model_cfg.get("INPUT_DIM", input_channels)
        self.cls_center_layers = make_fc_layers(self.model_cfg.CLS_FC, dim, num_class)
        self.box_center_layers = make_fc_layers(self.model_cfg.REG_FC, dim, self.box_coder.code_size)
        self.box_iou3d_layers = (make_fc_layers(self.model_cfg.IOU_FC, dim, 1)
                                 if self.model_cfg.get("IOU_FC", None) is not None else None)
        self.forward_ret_dict = {}
        self.fast_eval = True     # eval on CUDA: each Linear + BatchNorm1d + ReLU of the FC stacks as one tensor-core launch
        self._packed = None       # (fingerprint, {stack name: [(PackedLinear, epilogue, width)]})

    def generate_predicted_boxes(self, points, point_cls_preds, point_box_preds):
        pred_classes = point_cls_preds.max(dim=-1)[1]
        return point_cls_preds, self.box_coder.decode_torch(point_box_preds, points, pred_classes + 1)

    def get_loss(self, tb_dict=None):
        raise NotImplementedError("IASSD_Head target assignment / losses (pcdet/models/dense_heads/IASSD_head.py:169-1330) "
                                  "are outside the built hot path; train-mode forward returns the raw predictions")

    def _fc(self, name, feats):
        """One FC stack (point_head_template.py:36-47).  Eval on CUDA: BatchNorm folded into the Linear before it, each
        Linear + BN + ReLU one `tc_linear` launch (split-bf16 products, ~1e-5), the last Linear's width padded to a multiple
        of 4; the packed copies are rebuilt whenever a parameter or buffer of the head changes."""
        seq = getattr(self, name)
        if seq is None:
            return None
        if self.training or torch.is_grad_enabled() or not self.fast_eval or not feats.is_cuda:
            return seq(feats)
        from .pointnet2_modules import fold_conv_bn
        from .tc_linear import EPI_RELU, EPI_STORE, PackedLinear
        fp = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._packed is None or self._packed[0] != fp:
            self._packed = (fp, {})
        cache = self._packed[1]
        if name not in cache:
            layers, mods, k = [], list(seq), 0
            while k < len(mods):
                lin = mods[k]
                if k + 1 < len(mods) and isinstance(mods[k + 1], nn.BatchNorm1d):
                    w, b = fold_conv_bn(lin, mods[k + 1])
                    layers.append((PackedLinear(w, b, npass=2), EPI_RELU, w.shape[0]))
                    k += 3
                else:
                    w = lin.weight.detach().float()
                    b = lin.bias.detach().float() if lin.bias is not None else w.new_zeros(w.shape[0])
                    pad = (-w.shape[0]) % 4
                    if pad:
                        w, b = torch.cat([w, w.new_zeros(pad, w.shape[1])]), torch.cat([b, b.new_zeros(pad)])
                    layers.append((PackedLinear(w.contiguous(), b.contiguous(), npass=2), EPI_STORE, lin.weight.shape[0]))
                    k += 1
            cache[name] = layers
        x = feats if feats.stride(1) == 1 else feats.contiguous()
        for lin, epi, width in cache[name]:
            x = lin(x, epi)
        return x[:, :width]

    def forward(self, batch_dict):
        feats = batch_dict["centers_features"]
        coords = batch_dict["centers"]
        cls_preds = self._fc("cls_center_layers", feats)
        box_codes = self._fc("box_center_layers", feats)
        iou_preds = self._fc("box_iou3d_layers", feats)
        point_cls_preds, point_box_preds = self.generate_predicted_boxes(coords[:, 1:4], cls_preds, box_codes)
        batch_dict["batch_cls_preds"] = point_cls_preds
        batch_dict["batch_box_preds"] = point_box_preds
        batch_dict["box_iou3d_preds"] = iou_preds
        batch_dict["batch_index"] = coords[:, 0]
        batch_dict["cls_preds_normalized"] = False
        self.forward_ret_dict = {
            "center_cls_preds": cls_preds, "center_box_preds": box_codes, "ctr_offsets": batch_dict["ctr_offsets"],
            "centers": batch_dict["centers"], "centers_origin": batch_dict["centers_origin"],
            "sa_ins_preds": batch_dict["sa_ins_preds"], "sample_list_id": batch_dict["sample_list_id"],
            "box_iou3d_preds": iou_preds, "point_box_preds": point_box_preds,
        }
        return batch_dict

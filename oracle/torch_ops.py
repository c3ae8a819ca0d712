"""TEST INFRASTRUCTURE — CPU namespaces with the op API of pdanet_b200, backed by the C oracle.

`oracle.torch_ops` exposes the same names as `pdanet_b200.pointnet2_utils` (the reference's
`pointnet2_utils` API) on CPU tensors, and `oracle.torch_ops.nms_utils` the names of
`iou3d_nms_utils`.  Tests and bench.py's CPU baseline hand these to the model constructors
(`build_model(cfg, ops=..., nms_utils=...)`) to run the SAME module code without a GPU.  It has
no fused ops on purpose: the CPU baseline takes the reference's unfused route
(ball query -> group -> torch conv/BN/ReLU -> max-pool).
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

import oracle as _o


def furthest_point_sample(xyz, npoint):
    return _o.fps(xyz.contiguous(), npoint)


farthest_point_sample = furthest_point_sample


def furthest_point_sample_with_dist(dist, npoint):
    return _o.fps_with_dist(dist.contiguous(), npoint)


def gather_operation(features, idx):
    return _o.gather(features.contiguous(), idx.contiguous())


def grouping_operation(features, idx):
    return _o.group(features.contiguous(), idx.contiguous())


def ball_query(radius, nsample, xyz, new_xyz):
    return _o.ball_query(radius, nsample, xyz.contiguous(), new_xyz.contiguous())


def ball_query_dilated(max_radius, min_radius, nsample, xyz, new_xyz):
    return _o.ball_query_dilated(max_radius, min_radius, nsample, xyz.contiguous(), new_xyz.contiguous())


def topk_ctr_sample(cls_features, npoint):
    return _o.topk_ctr(cls_features.contiguous().float(), npoint)


class QueryAndGroup(nn.Module):
    """PB/pointnet2_utils.py:671-704."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            gf = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, gf], dim=1) if self.use_xyz else gf
        return grouped_xyz


class QueryDilatedAndGroup(nn.Module):
    """PB/pointnet2_utils.py:706-741."""

    def __init__(self, radius_in, radius_out, nsample, use_xyz=True):
        super().__init__()
        self.radius_in, self.radius_out, self.nsample, self.use_xyz = radius_in, radius_out, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query_dilated(self.radius_in, self.radius_out, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            gf = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, gf], dim=1) if self.use_xyz else gf
        return grouped_xyz


class QueryAndGroup_alone_grouped_density_directional(nn.Module):
    """PB/pointnet2_utils.py:557-614, statement by statement in torch."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        distances = torch.norm(grouped_xyz.permute(0, 2, 3, 1).contiguous() - new_xyz.unsqueeze(2), dim=-1)
        dens = torch.exp(-distances ** 2 / (2 * self.radius ** 2)) / (2.5 * self.radius)
        dens = dens.unsqueeze(-1).permute(0, 3, 1, 2).contiguous()
        direction = (grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)) / self.radius
        if features is not None:
            gf = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, dens, direction, gf], dim=1) if self.use_xyz else gf
        return grouped_xyz


class GroupAll(nn.Module):
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            gf = features.unsqueeze(2)
            return torch.cat([grouped_xyz, gf], dim=1) if self.use_xyz else gf
        return grouped_xyz


# ---- iou3d_nms_utils namespace -------------------------------------------------------------------

def _nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    order = scores.sort(dim=0, descending=True, stable=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = boxes[order].contiguous().float()
    keep = torch.zeros(b.shape[0], dtype=torch.int64)
    num = _o.nms_gpu(b, keep, thresh)
    return order[keep[:num]].contiguous(), None


def _nms_normal_gpu(boxes, scores, thresh, **kwargs):
    order = scores.sort(dim=0, descending=True, stable=True)[1]
    b = boxes[order].contiguous().float()
    keep = torch.zeros(b.shape[0], dtype=torch.int64)
    num = _o.nms_normal_gpu(b, keep, thresh)
    return order[keep[:num]].contiguous(), None


def _boxes_iou_bev(boxes_a, boxes_b):
    ans = torch.zeros(boxes_a.shape[0], boxes_b.shape[0])
    _o.boxes_iou_bev_cpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


nms_utils = types.SimpleNamespace(nms_gpu=_nms_gpu, nms_normal_gpu=_nms_normal_gpu, boxes_iou_bev=_boxes_iou_bev)

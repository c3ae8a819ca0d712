"""`torch.library` registration of the hot-path ops: `torch.ops.pdab.*`.

north_star asks for "a thin C-ABI torch custom-op layer".  The C ABI is include/pdab.h; this module registers its idx-
producing and gathering entry points as PyTorch custom operators (schema, CUDA implementation through ctypes -> libpdab.so,
fake/meta kernels for shape inference), so they show up in `torch.ops.pdab`, trace under `torch.export` /
`torch.compile(fullgraph=True)` without graph breaks and can be inspected by the dispatcher tooling.  The op API mirror
(`pointnet2_utils.py`) stays the primary surface — same names as the reference; these are the same kernels behind the
dispatcher.  Forward only where the reference's autograd Function returns None for the gradient (idx-producing ops,
PB/pointnet2_utils.py:31-33,251-253); gather / group register their backward through the C ABI's *_grad entry points.

CUDA only: there is no CPU implementation to dispatch to (a CPU tensor raises NotImplementedError from the dispatcher).
"""
from __future__ import annotations

import torch

from . import pointnet2_batch_cuda as _pn

_lib_def = torch.library.Library("pdab", "DEF")
_lib_def.define("furthest_point_sample(Tensor xyz, int npoint) -> Tensor")
_lib_def.define("furthest_point_sample_with_dist(Tensor dist, int npoint) -> Tensor")
_lib_def.define("ball_query(float radius, int nsample, Tensor xyz, Tensor new_xyz) -> Tensor")
_lib_def.define("gather_operation(Tensor features, Tensor idx) -> Tensor")
_lib_def.define("gather_operation_backward(Tensor grad_out, Tensor idx, int n) -> Tensor")
_lib_def.define("grouping_operation(Tensor features, Tensor idx) -> Tensor")
_lib_def.define("grouping_operation_backward(Tensor grad_out, Tensor idx, int n) -> Tensor")
_lib_def.define("nms_keep(Tensor sorted_boxes, float thresh) -> (Tensor, Tensor)")


def _fps(xyz, npoint):
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    _pn.farthest_point_sampling_wrapper(B, N, npoint, xyz.contiguous(), temp, idx)
    return idx


def _fps_dist(dist, npoint):
    B, N, _ = dist.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=dist.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=dist.device)
    _pn.furthest_point_sampling_with_dist_wrapper(B, N, npoint, dist.contiguous(), temp, idx)
    return idx


def _ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32, device=xyz.device)
    _pn.ball_query_wrapper(B, N, M, float(radius), nsample, new_xyz.contiguous(), xyz.contiguous(), idx)
    return idx


def _gather(features, idx):
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
    _pn.gather_points_wrapper(B, C, N, M, features.contiguous(), idx.contiguous(), out)
    return out


def _gather_bwd(grad_out, idx, n):
    B, C, M = grad_out.shape
    grad = torch.zeros(B, C, n, dtype=torch.float32, device=grad_out.device)
    _pn.gather_points_grad_wrapper(B, C, n, M, grad_out.contiguous(), idx.contiguous(), grad)
    return grad


def _group(features, idx):
    B, C, N = features.shape
    _, M, ns = idx.shape
    out = torch.empty(B, C, M, ns, dtype=torch.float32, device=features.device)
    _pn.group_points_wrapper(B, C, N, M, ns, features.contiguous(), idx.contiguous(), out)
    return out


def _group_bwd(grad_out, idx, n):
    B, C, M, ns = grad_out.shape
    grad = torch.zeros(B, C, n, dtype=torch.float32, device=grad_out.device)
    _pn.group_points_grad_wrapper(B, C, n, M, ns, grad_out.contiguous(), idx.contiguous(), grad)
    return grad


def _nms_keep(sorted_boxes, thresh):
    """Device-side rotated NMS of boxes already sorted by score: (keep positions (n) int64, num_keep (1) int32)."""
    from . import iou3d_nms_utils
    keep, num = iou3d_nms_utils.nms_batched(sorted_boxes.unsqueeze(0).contiguous(),
                                            torch.tensor([sorted_boxes.shape[0]], dtype=torch.int32,
                                                         device=sorted_boxes.device), float(thresh))
    return keep[0], num


for _name, _fn in (("furthest_point_sample", _fps), ("furthest_point_sample_with_dist", _fps_dist),
                   ("ball_query", _ball_query), ("gather_operation", _gather), ("gather_operation_backward", _gather_bwd),
                   ("grouping_operation", _group), ("grouping_operation_backward", _group_bwd), ("nms_keep", _nms_keep)):
    _lib_def.impl(_name, _fn, "CUDA")


# fake (meta) kernels: shapes and dtypes only
@torch.library.register_fake("pdab::furthest_point_sample")
def _(xyz, npoint):
    return xyz.new_empty((xyz.shape[0], npoint), dtype=torch.int32)


@torch.library.register_fake("pdab::furthest_point_sample_with_dist")
def _(dist, npoint):
    return dist.new_empty((dist.shape[0], npoint), dtype=torch.int32)


@torch.library.register_fake("pdab::ball_query")
def _(radius, nsample, xyz, new_xyz):
    return xyz.new_empty((xyz.shape[0], new_xyz.shape[1], nsample), dtype=torch.int32)


@torch.library.register_fake("pdab::gather_operation")
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1]))


@torch.library.register_fake("pdab::gather_operation_backward")
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


@torch.library.register_fake("pdab::grouping_operation")
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1], idx.shape[2]))


@torch.library.register_fake("pdab::grouping_operation_backward")
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


@torch.library.register_fake("pdab::nms_keep")
def _(sorted_boxes, thresh):
    return (sorted_boxes.new_empty((sorted_boxes.shape[0],), dtype=torch.int64),
            sorted_boxes.new_empty((1,), dtype=torch.int32))


# autograd: the gradient of a gather / group w.r.t. the features is a scatter-add through the same indices
def _gather_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _gather_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    return torch.ops.pdab.gather_operation_backward(grad.contiguous(), idx, ctx.n), None


def _group_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    return torch.ops.pdab.grouping_operation_backward(grad.contiguous(), idx, ctx.n), None


torch.library.register_autograd("pdab::gather_operation", _gather_backward, setup_context=_gather_setup)
torch.library.register_autograd("pdab::grouping_operation", _group_backward, setup_context=_gather_setup)

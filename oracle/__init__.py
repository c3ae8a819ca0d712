"""TEST INFRASTRUCTURE — Python face of the CPU oracle (oracle/pdab_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this package.  pdanet_b200 (the product) never does.

Functions take and return CPU torch tensors (or numpy arrays) with the same
layouts, dtypes and caller-allocates conventions as the reference pybind
module `pointnet2_batch_cuda` (PB/src/pointnet2_api.cpp:12-33) and
`iou3d_nms_cuda` (IOU/src/iou3d_nms_api.cpp:11-17).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libpdab_oracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "pdab_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libpdab_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_box_overlap.restype = C.c_float
        _lib.orc_iou_bev.restype = C.c_float
        _lib.orc_nms.restype = C.c_int
        _lib.orc_nms_normal.restype = C.c_int
        _lib.orc_opt_n_threads.restype = C.c_int
    return _lib


def _np(t, dtype):
    if isinstance(t, torch.Tensor):
        assert t.device.type == "cpu", "oracle works on CPU tensors"
        assert t.is_contiguous()
        a = t.detach().numpy()
    else:
        a = t
    assert a.dtype == dtype, (a.dtype, dtype)
    assert a.flags["C_CONTIGUOUS"]
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(t):
    return _p(_np(t, np.float32))


def _i(t):
    return _p(_np(t, np.int32))


# ---------------------------------------------------------------- raw ops (pybind arity)

def farthest_point_sampling_wrapper(b, n, m, xyz, temp, idx):
    lib().orc_fps(C.c_int(b), C.c_int(n), C.c_int(m), _f(xyz), _f(temp), _i(idx))
    return 1


def furthest_point_sampling_with_dist_wrapper(b, n, m, dist, temp, idx):
    lib().orc_fps_with_dist(C.c_int(b), C.c_int(n), C.c_int(m), _f(dist), _f(temp), _i(idx))
    return 2


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    lib().orc_gather(C.c_int(b), C.c_int(c), C.c_int(n), C.c_int(npoints), _f(points), _i(idx), _f(out))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    lib().orc_gather_grad(C.c_int(b), C.c_int(c), C.c_int(n), C.c_int(npoints), _f(grad_out), _i(idx), _f(grad_points))
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    lib().orc_ball_query(C.c_int(b), C.c_int(n), C.c_int(m), C.c_float(radius), C.c_int(nsample),
                         _f(new_xyz), _f(xyz), _i(idx))
    return 1


def ball_query_dilated_wrapper(b, n, m, max_radius, min_radius, nsample, new_xyz, xyz, idx):
    lib().orc_ball_query_dilated(C.c_int(b), C.c_int(n), C.c_int(m), C.c_float(max_radius), C.c_float(min_radius),
                                 C.c_int(nsample), _f(new_xyz), _f(xyz), _i(idx))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    lib().orc_group(C.c_int(b), C.c_int(c), C.c_int(n), C.c_int(npoints), C.c_int(nsample), _f(points), _i(idx), _f(out))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    lib().orc_group_grad(C.c_int(b), C.c_int(c), C.c_int(n), C.c_int(npoints), C.c_int(nsample),
                         _f(grad_out), _i(idx), _f(grad_points))
    return 1


def nms_gpu(boxes, keep, thresh):
    """Same contract as the reference pybind nms_gpu (IOU/src/iou3d_nms.cpp:90): boxes (n,7)
    sorted by score, keep (n) int64 CPU, returns num_to_keep."""
    n = boxes.shape[0]
    return lib().orc_nms(_f(boxes), C.c_int(n), C.c_float(thresh), _p(_np(keep, np.int64)))


def nms_normal_gpu(boxes, keep, thresh):
    n = boxes.shape[0]
    return lib().orc_nms_normal(_f(boxes), C.c_int(n), C.c_float(thresh), _p(_np(keep, np.int64)))


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans):
    lib().orc_boxes_iou_bev(C.c_int(boxes_a.shape[0]), _f(boxes_a), C.c_int(boxes_b.shape[0]), _f(boxes_b), _f(ans))
    return 1


boxes_iou_bev_gpu = boxes_iou_bev_cpu


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans):
    lib().orc_boxes_overlap_bev(C.c_int(boxes_a.shape[0]), _f(boxes_a), C.c_int(boxes_b.shape[0]), _f(boxes_b), _f(ans))
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    lib().orc_three_nn(C.c_int(b), C.c_int(n), C.c_int(m), _f(unknown), _f(known), _f(dist2), _i(idx))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    lib().orc_three_interpolate(C.c_int(b), C.c_int(c), C.c_int(m), C.c_int(n), _f(points), _i(idx), _f(weight), _f(out))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    lib().orc_three_interpolate_grad(C.c_int(b), C.c_int(c), C.c_int(n), C.c_int(m), _f(grad_out), _i(idx), _f(weight),
                                     _f(grad_points))


def points_in_boxes(points: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """points (B,M,3), boxes (B,T,7) -> (B,M) int32, -1 = background (roiaware_pool3d_utils.py:28-41)."""
    B, M, _ = points.shape
    out = torch.full((B, M), -1, dtype=torch.int32)
    lib().orc_points_in_boxes(C.c_int(B), C.c_int(boxes.shape[1]), C.c_int(M), _f(boxes.contiguous()),
                              _f(points.contiguous()), _i(out))
    return out


def three_nn(unknown: torch.Tensor, known: torch.Tensor):
    B, N, _ = unknown.shape
    dist2 = torch.zeros(B, N, 3, dtype=torch.float32)
    idx = torch.zeros(B, N, 3, dtype=torch.int32)
    three_nn_wrapper(B, N, known.shape[1], unknown.contiguous(), known.contiguous(), dist2, idx)
    return dist2, idx


def three_interpolate(features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    B, c, m = features.shape
    n = idx.shape[1]
    out = torch.zeros(B, c, n, dtype=torch.float32)
    three_interpolate_wrapper(B, c, m, n, features.contiguous(), idx.contiguous(), weight.contiguous(), out)
    return out


# ---------------------------------------------------------------- convenience (allocating) forms

def fps(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    B, N, _ = xyz.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    farthest_point_sampling_wrapper(B, N, npoint, xyz.contiguous(), temp, idx)
    return idx


def fps_with_dist(dist: torch.Tensor, npoint: int) -> torch.Tensor:
    B, N, _ = dist.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    furthest_point_sampling_with_dist_wrapper(B, N, npoint, dist.contiguous(), temp, idx)
    return idx


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32)
    ball_query_wrapper(B, N, M, radius, nsample, new_xyz.contiguous(), xyz.contiguous(), idx)
    return idx


def ball_query_dilated(max_radius, min_radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32)
    ball_query_dilated_wrapper(B, N, M, max_radius, min_radius, nsample, new_xyz.contiguous(), xyz.contiguous(), idx)
    return idx


def gather(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    B, Cc, N = features.shape
    out = torch.empty(B, Cc, idx.shape[1], dtype=torch.float32)
    gather_points_wrapper(B, Cc, N, idx.shape[1], features.contiguous(), idx.contiguous(), out)
    return out


def group(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    B, Cc, N = features.shape
    _, M, ns = idx.shape
    out = torch.empty(B, Cc, M, ns, dtype=torch.float32)
    group_points_wrapper(B, Cc, N, M, ns, features.contiguous(), idx.contiguous(), out)
    return out


def topk_ctr(cls_features: torch.Tensor, npoint: int) -> torch.Tensor:
    B, N, Cc = cls_features.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    lib().orc_topk_ctr(C.c_int(B), C.c_int(N), C.c_int(Cc), C.c_int(npoint), _f(cls_features.contiguous()), _i(idx))
    return idx


def pda_group(radius, nsample, xyz, new_xyz, features):
    """PDA grouper (PB/pointnet2_utils.py:557-614): returns (out (B,7+C,M,ns), idx)."""
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    Cc = features.shape[1]
    idx = ball_query(radius, nsample, xyz, new_xyz)
    out = torch.empty(B, 7 + Cc, M, nsample, dtype=torch.float32)
    lib().orc_pda_group(C.c_int(B), C.c_int(Cc), C.c_int(N), C.c_int(M), C.c_int(nsample), C.c_float(radius),
                        _f(xyz.contiguous()), _f(new_xyz.contiguous()), _f(features.contiguous()), _i(idx), _f(out))
    return out, idx


def sa_mlp_maxpool(radius, nsample, xyz, new_xyz, features, weights, biases):
    """Fused plain-SA scale in fp32 (ball query -> group -> folded MLP -> max-pool).
    weights[l]: (cout_l, cin_l); biases[l]: (cout_l).  Returns (B, cout_last, M)."""
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    Cc = 0 if features is None else features.shape[1]
    idx = ball_query(radius, nsample, xyz, new_xyz)
    dims = [3 + Cc] + [int(w.shape[0]) for w in weights]
    L = len(weights)
    ws = [np.ascontiguousarray(w.detach().numpy(), dtype=np.float32) for w in weights]
    bs = [np.ascontiguousarray(x.detach().numpy(), dtype=np.float32) for x in biases]
    wp = (C.c_void_p * L)(*[w.ctypes.data for w in ws])
    bp = (C.c_void_p * L)(*[x.ctypes.data for x in bs])
    dims_a = (C.c_int * (L + 1))(*dims)
    out = torch.empty(B, dims[-1], M, dtype=torch.float32)
    feats = features.contiguous() if features is not None else torch.zeros(1)
    lib().orc_sa_mlp_maxpool(C.c_int(B), C.c_int(Cc), C.c_int(N), C.c_int(M), C.c_int(nsample),
                             _f(xyz.contiguous()), _f(new_xyz.contiguous()), _f(feats), _i(idx),
                             C.c_int(L), dims_a, wp, bp, _f(out))
    return out


def nms(boxes: torch.Tensor, scores: torch.Tensor, thresh: float):
    """iou3d_nms_utils.nms_gpu semantics (IOU/iou3d_nms_utils.py:84-99) on CPU tensors."""
    order = scores.sort(dim=0, descending=True, stable=True)[1]
    b = boxes[order].contiguous()
    keep = torch.zeros(b.shape[0], dtype=torch.int64)
    num = nms_gpu(b, keep, thresh)
    return order[keep[:num]].contiguous()

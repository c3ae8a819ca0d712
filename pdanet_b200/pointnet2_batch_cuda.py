"""Drop-in for the reference pybind11 module `pointnet2_batch_cuda`.

Same function names, positional arities, tensor layouts and caller-allocates
convention as PB/src/pointnet2_api.cpp:12-33 (PB = pcdet/ops/pointnet2/
pointnet2_batch in the reference), so the reference's `pointnet2_utils.py` can
`import pointnet2_batch_cuda as pointnet2` from here unchanged.  Every call goes
straight to libpdab.so (include/pdab.h) on the tensors' device and the current
torch CUDA stream.  Differences from the reference, all deliberate:
  * wrong device / dtype / non-contiguous input raises RuntimeError instead of
    fprintf + exit(-1) (PB/src/ball_query.cpp:17-29);
  * launches go to torch's current stream with a device guard, not to the
    legacy default stream of whatever device happens to be current.
Functions of the reference module that PDA-SSD never reaches (ellipsoid_query,
chamfer) are not provided and raise NotImplementedError naming SURVEY.md section 8(f).
"""
from __future__ import annotations

import torch

from . import _lib


def _chk(t: torch.Tensor, name: str, dtype: torch.dtype, shape=None) -> int:
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _same_device(*ts):
    d = ts[0].device
    for t in ts[1:]:
        if t.device != d:
            raise RuntimeError("all tensors must be on the same CUDA device")
    return torch.cuda.device(d)


def farthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    """PB/src/sampling.cpp:34-43."""
    px = _chk(points_tensor, "xyz", torch.float32, (b, n, 3))
    pt = _chk(temp_tensor, "temp", torch.float32, (b, n))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, m))
    with _same_device(points_tensor, temp_tensor, idx_tensor):
        _lib.call("pdab_fps", b, n, m, px, pt, pi, _stream(points_tensor))
    return 1


def furthest_point_sampling_with_dist_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    """PB/src/sampling.cpp:46-56 (returns 2 there)."""
    px = _chk(points_tensor, "dist", torch.float32, (b, n, n))
    pt = _chk(temp_tensor, "temp", torch.float32, (b, n))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, m))
    with _same_device(points_tensor, temp_tensor, idx_tensor):
        _lib.call("pdab_fps_with_dist", b, n, m, px, pt, pi, _stream(points_tensor))
    return 2


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    """PB/src/sampling.cpp:11-19."""
    pp = _chk(points_tensor, "points", torch.float32, (b, c, n))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, npoints))
    po = _chk(out_tensor, "out", torch.float32, (b, c, npoints))
    with _same_device(points_tensor, idx_tensor, out_tensor):
        _lib.call("pdab_gather_points", b, c, n, npoints, pp, pi, po, _stream(points_tensor))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    """PB/src/sampling.cpp:22-31."""
    pg = _chk(grad_out_tensor, "grad_out", torch.float32, (b, c, npoints))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, npoints))
    po = _chk(grad_points_tensor, "grad_points", torch.float32, (b, c, n))
    with _same_device(grad_out_tensor, idx_tensor, grad_points_tensor):
        _lib.call("pdab_gather_points_grad", b, c, n, npoints, pg, pi, po, _stream(grad_out_tensor))
    return 1


# Clouds of at least this many points are searched through a hashed cell list (pdab_ball_query_grid: 27 cells per centre instead of
# the whole cloud, dense balls fall back to the in-order scan inside the kernel; the same idx bit for bit); smaller ones are scanned
# in order with early exit (building the list costs more than it saves).  None disables the cell list.
CELL_LIST_MIN_POINTS = 8192


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    """PB/src/ball_query.cpp:32-43.  idx must be pre-zeroed by the caller."""
    pn = _chk(new_xyz_tensor, "new_xyz", torch.float32, (b, m, 3))
    px = _chk(xyz_tensor, "xyz", torch.float32, (b, n, 3))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, m, nsample))
    with _same_device(new_xyz_tensor, xyz_tensor, idx_tensor):
        if CELL_LIST_MIN_POINTS is not None and n >= CELL_LIST_MIN_POINTS and nsample <= 256 and radius > 0 and b > 0 and m > 0:
            ws = torch.empty(_lib.lib().pdab_ball_query_grid_workspace_bytes(b, n, m), dtype=torch.uint8,
                             device=xyz_tensor.device)
            _lib.call("pdab_ball_query_grid", b, n, m, float(radius), nsample, pn, px, pi, ws.data_ptr(), _stream(xyz_tensor))
        else:
            _lib.call("pdab_ball_query", b, n, m, float(radius), nsample, pn, px, pi, _stream(xyz_tensor))
    return 1


def ball_query_dilated_wrapper(b, n, m, max_radius, min_radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    """PB/src/ball_query.cpp:45-56."""
    pn = _chk(new_xyz_tensor, "new_xyz", torch.float32, (b, m, 3))
    px = _chk(xyz_tensor, "xyz", torch.float32, (b, n, 3))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, m, nsample))
    with _same_device(new_xyz_tensor, xyz_tensor, idx_tensor):
        _lib.call("pdab_ball_query_dilated", b, n, m, float(max_radius), float(min_radius), nsample, pn, px, pi,
                  _stream(xyz_tensor))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    """PB/src/group_points.cpp:29-39."""
    pp = _chk(points_tensor, "points", torch.float32, (b, c, n))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, npoints, nsample))
    po = _chk(out_tensor, "out", torch.float32, (b, c, npoints, nsample))
    with _same_device(points_tensor, idx_tensor, out_tensor):
        _lib.call("pdab_group_points", b, c, n, npoints, nsample, pp, pi, po, _stream(points_tensor))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    """PB/src/group_points.cpp:16-27."""
    pg = _chk(grad_out_tensor, "grad_out", torch.float32, (b, c, npoints, nsample))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, npoints, nsample))
    po = _chk(grad_points_tensor, "grad_points", torch.float32, (b, c, n))
    with _same_device(grad_out_tensor, idx_tensor, grad_points_tensor):
        _lib.call("pdab_group_points_grad", b, c, n, npoints, nsample, pg, pi, po, _stream(grad_out_tensor))
    return 1


def _out_of_scope(name):
    def fn(*_a, **_k):
        raise NotImplementedError(
            f"{name} is outside the PDA-SSD hot path (SURVEY.md section 8(f), 'next'); not built in this round")
    fn.__name__ = name
    return fn


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    """PB/src/interpolate.cpp:21-33."""
    pu = _chk(unknown_tensor, "unknown", torch.float32, (b, n, 3))
    pk = _chk(known_tensor, "known", torch.float32, (b, m, 3))
    pd = _chk(dist2_tensor, "dist2", torch.float32, (b, n, 3))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, n, 3))
    with _same_device(unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
        _lib.call("pdab_three_nn", b, n, m, pu, pk, pd, pi, _stream(unknown_tensor))


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    """PB/src/interpolate.cpp:36-47."""
    pp = _chk(points_tensor, "points", torch.float32, (b, c, m))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, n, 3))
    pw = _chk(weight_tensor, "weight", torch.float32, (b, n, 3))
    po = _chk(out_tensor, "out", torch.float32, (b, c, n))
    with _same_device(points_tensor, idx_tensor, weight_tensor, out_tensor):
        _lib.call("pdab_three_interpolate", b, c, m, n, pp, pi, pw, po, _stream(points_tensor))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    """PB/src/interpolate.cpp:50-61."""
    pg = _chk(grad_out_tensor, "grad_out", torch.float32, (b, c, n))
    pi = _chk(idx_tensor, "idx", torch.int32, (b, n, 3))
    pw = _chk(weight_tensor, "weight", torch.float32, (b, n, 3))
    po = _chk(grad_points_tensor, "grad_points", torch.float32, (b, c, m))
    with _same_device(grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
        _lib.call("pdab_three_interpolate_grad", b, c, n, m, pg, pi, pw, po, _stream(grad_out_tensor))


ellipsoid_query = _out_of_scope("ellipsoid_query")
chamfer_forward = _out_of_scope("chamfer_forward")
chamfer_backward = _out_of_scope("chamfer_backward")

"""Host-side mirror of the reference's `pointnet2_utils` API over libpdab.so.

Keeps the reference's public surface — furthest_point_sample, furthest_point_sample_with_dist,
gather_operation, grouping_operation, ball_query, ball_query_dilated, QueryAndGroup,
QueryAndGroup_alone_grouped_density_directional, QueryDilatedAndGroup, GroupAll
(PB/pointnet2_utils.py:36,65,101,225,256,287,671,557,706,743) — with the same argument
order, tensor layouts, dtypes (int32 indices) and differentiability, and adds the fused
entry points that have no reference counterpart (topk_ctr_sample, pda_group, sa_fused).

Every op runs on the CUDA device of its inputs through `pointnet2_batch_cuda`
(our C-ABI shim); CPU tensors are rejected — there is no fallback path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from . import pointnet2_batch_cuda as pointnet2


class FarthestPointSampling(Function):
    """xyz (B,N,3) fp32 -> idx (B,npoint) int32.  PB/pointnet2_utils.py:10-33."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2.farthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


class FurthestPointSamplingWithDist(Function):
    """dist (B,N,N) fp32 -> idx (B,npoint) int32.  PB/pointnet2_utils.py:39-62."""

    @staticmethod
    def forward(ctx, dist: torch.Tensor, npoint: int) -> torch.Tensor:
        assert dist.is_contiguous()
        B, N, _ = dist.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=dist.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=dist.device)
        pointnet2.furthest_point_sampling_with_dist_wrapper(B, N, npoint, dist, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample_with_dist = FurthestPointSamplingWithDist.apply


# Gradients of gather / group / three_interpolate: a sorted segmented sum (bit-identical from run to run) by default; False =
# the reference's float atomicAdd scatter (PB/src/group_points_gpu.cu:30), faster for a handful of channels but its
# summation order — and with it the last bits of every gradient — changes between runs.
DETERMINISTIC_BACKWARD = True


def _segment_sum_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int, weight: Optional[torch.Tensor] = None, div: int = 1):
    """grad_out (B, C, E / div), idx (B, E) int: which of the n points each slot read -> (B, C, n), deterministic."""
    B, C = grad_out.shape[0], grad_out.shape[1]
    flat = idx.reshape(B, -1)
    E = flat.shape[1]
    vals, order = torch.sort(flat.long(), dim=1, stable=True)
    bounds = torch.arange(n + 1, device=idx.device, dtype=torch.long).unsqueeze(0).expand(B, -1).contiguous()
    seg = torch.searchsorted(vals.contiguous(), bounds).int().contiguous()
    order = order.int().contiguous()
    grad = torch.empty(B, C, n, dtype=torch.float32, device=grad_out.device)
    g = grad_out.reshape(B, C, E // div).contiguous()
    with torch.cuda.device(grad_out.device):
        _lib.call("pdab_segment_sum_grad", B, C, n, E, div, g.data_ptr(),
                  None if weight is None else weight.reshape(B, E).contiguous().data_ptr(), order.data_ptr(),
                  seg.data_ptr(), grad.data_ptr(), torch.cuda.current_stream(grad_out.device).cuda_stream)
    return grad


class GatherOperation(Function):
    """features (B,C,N), idx (B,npoint) -> (B,C,npoint).  PB/pointnet2_utils.py:67-98."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, npoint, dtype=torch.float32, device=features.device)
        pointnet2.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        if DETERMINISTIC_BACKWARD and grad_out.is_cuda:
            return _segment_sum_grad(grad_out, idx, N), None
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """unknown (B,N,3), known (B,M,3) -> (dist (B,N,3) L2 distances, idx (B,N,3) int32).  PB/pointnet2_utils.py:104-133."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor):
        assert unknown.is_contiguous() and known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=unknown.device)
        pointnet2.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,C,M), idx (B,n,3), weight (B,n,3) -> (B,C,n).  PB/pointnet2_utils.py:136-181."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous() and idx.is_contiguous() and weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        pointnet2.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        if DETERMINISTIC_BACKWARD and grad_out.is_cuda:
            return _segment_sum_grad(grad_out, idx, m, weight=weight, div=3), None, None
        grad_features = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        pointnet2.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample).  PB/pointnet2_utils.py:184-222."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, nfeatures, nsample, dtype=torch.float32, device=features.device)
        pointnet2.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        if DETERMINISTIC_BACKWARD and grad_out.is_cuda:
            return _segment_sum_grad(grad_out, idx, N), None
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """(radius, nsample, xyz (B,N,3), new_xyz (B,M,3)) -> idx (B,M,nsample) int32.  PB/pointnet2_utils.py:228-253."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
        pointnet2.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class BallQueryDilated(Function):
    """PB/pointnet2_utils.py:258-284."""

    @staticmethod
    def forward(ctx, max_radius: float, min_radius: float, nsample: int, xyz: torch.Tensor,
                new_xyz: torch.Tensor) -> torch.Tensor:
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
        pointnet2.ball_query_dilated_wrapper(B, N, npoint, max_radius, min_radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None, None


ball_query_dilated = BallQueryDilated.apply


# --------------------------------------------------------------------------- fused entry points

def topk_ctr_sample(cls_features: torch.Tensor, npoint: int) -> torch.Tensor:
    """Class-aware sampling: cls_features (B,N,num_class) -> idx (B,npoint) int32 ordered by
    (max-class logit desc, index asc).  One kernel for cls.max(-1) -> sigmoid -> topk -> .int()
    (PB/pointnet2_modules.py:761-770)."""
    if not cls_features.is_cuda:
        raise RuntimeError("cls_features must be a CUDA tensor")
    cls_features = cls_features.contiguous().float()
    B, N, C = cls_features.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=cls_features.device)
    with torch.cuda.device(cls_features.device):
        _lib.call("pdab_topk_ctr", B, N, C, npoint, cls_features.data_ptr(), idx.data_ptr(),
                  torch.cuda.current_stream(cls_features.device).cuda_stream)
    return idx


def pda_group(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor,
              return_idx: bool = False):
    """Fused PDA grouper (forward only): (B, 7+C, M, nsample) with channels
    [xyz, density, direction, features] — PB/pointnet2_utils.py:567-614 in one kernel."""
    for t in (xyz, new_xyz, features):
        if not t.is_cuda:
            raise RuntimeError("pda_group needs CUDA tensors")
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and features.is_contiguous()
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = features.shape[1]
    out = torch.empty(B, 7 + C, M, nsample, dtype=torch.float32, device=xyz.device)
    idx = torch.empty(B, M, nsample, dtype=torch.int32, device=xyz.device) if return_idx else None
    with torch.cuda.device(xyz.device):
        _lib.call("pdab_pda_group", B, C, N, M, float(radius), nsample, xyz.data_ptr(), new_xyz.data_ptr(),
                  features.data_ptr(), out.data_ptr(), idx.data_ptr() if return_idx else None,
                  torch.cuda.current_stream(xyz.device).cuda_stream)
    return (out, idx) if return_idx else out


def pda_group_tokens(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor,
                     features_t: torch.Tensor, return_idx: bool = False):
    """Token-major fused PDA grouper (forward only).  features_t (B,N,C) point-major.  Returns
    X (B, M, nsample, 8+C): [xyz(3), density, direction(3), 0, features(C)] per (centre, neighbour) token."""
    for t in (xyz, new_xyz, features_t):
        if not t.is_cuda:
            raise RuntimeError("pda_group_tokens needs CUDA tensors")
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and features_t.is_contiguous()
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = features_t.shape[2]
    pitch = 8 + C
    out = torch.empty(B, M, nsample, pitch, dtype=torch.float32, device=xyz.device)
    idx = torch.empty(B, M, nsample, dtype=torch.int32, device=xyz.device) if return_idx else None
    with torch.cuda.device(xyz.device):
        _lib.call("pdab_pda_group_tokens", B, C, N, M, float(radius), nsample, pitch, xyz.data_ptr(),
                  new_xyz.data_ptr(), features_t.data_ptr(), out.data_ptr(), idx.data_ptr() if return_idx else None,
                  torch.cuda.current_stream(xyz.device).cuda_stream)
    return (out, idx) if return_idx else out


def pda_encode_supported(c: int, nsample: int) -> bool:
    """Shapes the fused token encoder covers (csrc/pda_encode.cu)."""
    return c in (64, 128) and nsample in (16, 32)


def pda_encode_params(w1, b1, w2, b2, dens, gamma, beta) -> torch.Tensor:
    """Pack the parameters of `pda_encode_ln` (layout in include/pdab.h): position MLP (w1 (C/2,12), b1, w2 (C,C/2), b2),
    DensityNet [(w (16,1), b), (w (8,16), b), (w (1,8), b)], LayerNorm gamma / beta (4C)."""
    C = w2.shape[0]
    assert w1.shape == (C // 2, 12) and w2.shape == (C, C // 2) and gamma.numel() == 4 * C
    (dw1, db1), (dw2, db2), (dw3, db3) = dens
    assert dw1.shape == (16, 1) and dw2.shape == (8, 16) and dw3.shape == (1, 8)
    parts = [w1.reshape(-1), b1, w2.t().contiguous().reshape(-1), b2, dw1.reshape(-1), db1, dw2.reshape(-1), db2,
             dw3.reshape(-1), db3.reshape(-1), torch.zeros(3, device=w1.device), gamma, beta]
    out = torch.cat([q.detach().float().reshape(-1) for q in parts]).contiguous()
    assert out.numel() == _lib.lib().pdab_pda_encode_param_floats(C)
    return out


def pda_encode_ln(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor, features_t: torch.Tensor,
                  glob: torch.Tensor, params: torch.Tensor, eps: float, split_half: bool = False):
    """Fused PDA token encoder (forward only): LayerNorm(cat[pos, feat*scale, feat, glob]) per (centre, neighbour)
    token, (B*M*nsample, 4C); see include/pdab.h `pdab_pda_encode_ln`.  split_half: the rows are returned as a
    `tc_linear.SplitHalf` — a (hi, lo) pair of fp16 planes, hi + lo = the fp32 row to ~2^-22 (`pdab_pda_encode_ln_h`)."""
    for t in (xyz, new_xyz, features_t, glob, params):
        if not t.is_cuda:
            raise RuntimeError("pda_encode_ln needs CUDA tensors")
        assert t.is_contiguous() and t.dtype == torch.float32
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = features_t.shape[2]
    assert glob.shape == (B * M, C)
    if split_half:
        from .tc_linear import SplitHalf
        y = SplitHalf.empty(B * M * nsample, 4 * C, xyz.device)
        with torch.cuda.device(xyz.device):
            _lib.call("pdab_pda_encode_ln_h", B, C, N, M, float(radius), nsample, xyz.data_ptr(), new_xyz.data_ptr(),
                      features_t.data_ptr(), glob.data_ptr(), params.data_ptr(), float(eps), y.hi.data_ptr(),
                      y.lo.data_ptr(), torch.cuda.current_stream(xyz.device).cuda_stream)
        return y
    y = torch.empty(B * M * nsample, 4 * C, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.call("pdab_pda_encode_ln", B, C, N, M, float(radius), nsample, xyz.data_ptr(), new_xyz.data_ptr(),
                  features_t.data_ptr(), glob.data_ptr(), params.data_ptr(), float(eps), y.data_ptr(),
                  torch.cuda.current_stream(xyz.device).cuda_stream)
    return y


def _stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def pda_assemble_ln_split(pos, X, scale, glob, nsample, norm):
    """(hi, lo) of LayerNorm(cat[pos, feat*scale, feat, glob]) per token; see include/pdab.h."""
    T, C = pos.shape
    assert pos.is_contiguous() and X.is_contiguous() and scale.is_contiguous() and glob.is_contiguous()
    hi = torch.empty(T, 4 * C, dtype=torch.float32, device=pos.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(pos.device):
        _lib.call("pdab_pda_assemble_ln_split", T, nsample, C, X.shape[-1], pos.data_ptr(), X.data_ptr(),
                  scale.data_ptr(), glob.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(), float(norm.eps),
                  hi.data_ptr(), lo.data_ptr(), _stream_of(pos))
    return hi, lo


def pda_assemble_ln(pos, X, scale, glob, nsample, norm):
    """LayerNorm(cat[pos, feat*scale, feat, glob]) per token, un-split (T, 4C); see include/pdab.h."""
    T, C = pos.shape
    assert pos.is_contiguous() and X.is_contiguous() and scale.is_contiguous() and glob.is_contiguous()
    y = torch.empty(T, 4 * C, dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        _lib.call("pdab_pda_assemble_ln_split", T, nsample, C, X.shape[-1], pos.data_ptr(), X.data_ptr(),
                  scale.data_ptr(), glob.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(), float(norm.eps),
                  y.data_ptr(), None, _stream_of(pos))
    return y


def group_attention(qkv, nsample, heads, npass: int = 3):
    """softmax(q k^T / sqrt(hd)) v inside each neighbourhood of `nsample` consecutive tokens; qkv (T, 3E) -> (T, E).
    npass: 3 = 3xTF32 products (fp32-level), 2 = split-bf16 products (the class of the split-bf16 GEMMs)."""
    if not qkv.is_cuda:
        raise RuntimeError("group_attention needs CUDA tensors")
    T, E3 = qkv.shape
    E = E3 // 3
    assert qkv.is_contiguous() and T % nsample == 0 and E % heads == 0
    ctx = torch.empty(T, E, dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.call("pdab_group_attention", T // nsample, nsample, heads, E // heads, int(npass), qkv.data_ptr(),
                  ctx.data_ptr(), _stream_of(qkv))
    return ctx


def group_attention_h(qkv, nsample, heads):
    """fp16 form of `group_attention`: qkv (T, 3E) fp16 -> ctx (T, E) fp16 (fp16 MMAs, fp32 accumulation and softmax)."""
    if not qkv.is_cuda:
        raise RuntimeError("group_attention_h needs CUDA tensors")
    T, E3 = qkv.shape
    E = E3 // 3
    assert qkv.dtype == torch.float16 and qkv.is_contiguous() and T % nsample == 0 and E % heads == 0
    ctx = torch.empty(T, E, dtype=torch.float16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.call("pdab_group_attention_h", T // nsample, nsample, heads, E // heads, qkv.data_ptr(), ctx.data_ptr(),
                  _stream_of(qkv))
    return ctx


def sa_fused_supported(c0: int, dims: Sequence[int], nsample: int) -> bool:
    """Shapes the fused plain-SA kernel covers (see csrc/sa_fused.cu)."""
    return len(dims) == 3 and nsample <= 32 and c0 <= 8 and tuple(dims) in ((16, 16, 32), (32, 32, 64))


def sa_fused(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor,
             features: Optional[torch.Tensor], weights: List[torch.Tensor], biases: List[torch.Tensor]) -> torch.Tensor:
    """Fused plain-SA scale (forward only): ball query -> group -> folded (conv,BN,ReLU) x L -> max-pool.
    weights[l] (cout_l, cin_l), biases[l] (cout_l) are the BN-folded parameters.  Returns (B, cout_last, M)."""
    assert xyz.is_cuda and xyz.is_contiguous() and new_xyz.is_contiguous()
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = 0 if features is None else features.shape[1]
    if features is not None:
        assert features.is_contiguous()
    L = len(weights)
    dims = [3 + C] + [int(w.shape[0]) for w in weights]
    for l, (w, bvec) in enumerate(zip(weights, biases)):
        assert w.is_cuda and w.is_contiguous() and w.dtype == torch.float32 and tuple(w.shape) == (dims[l + 1], dims[l])
        assert bvec.is_cuda and bvec.is_contiguous() and bvec.dtype == torch.float32 and bvec.numel() == dims[l + 1]
    out = torch.empty(B, dims[-1], M, dtype=torch.float32, device=xyz.device)
    dims_a = (ctypes.c_int * (L + 1))(*dims)
    w_a = (ctypes.c_void_p * L)(*[w.data_ptr() for w in weights])
    b_a = (ctypes.c_void_p * L)(*[x.data_ptr() for x in biases])
    with torch.cuda.device(xyz.device):
        _lib.call("pdab_sa_fused", B, C, N, M, float(radius), nsample, xyz.data_ptr(), new_xyz.data_ptr(),
                  features.data_ptr() if features is not None else None, L, dims_a, w_a, b_a, out.data_ptr(),
                  torch.cuda.current_stream(xyz.device).cuda_stream)
    return out


def sa_fused_pair_supported(c0: int, dims_a: Sequence[int], ns_a: int, dims_b: Sequence[int], ns_b: int) -> bool:
    """Two-scale shapes the pair kernel covers (csrc/sa_fused.cu: one scan of the cloud for both radii)."""
    return (c0 <= 8 and tuple(dims_a) == (16, 16, 32) and tuple(dims_b) == (32, 32, 64) and ns_a <= 32 and ns_b <= 32)


def sa_fused_pair(radii, nsamples, xyz, new_xyz, features, weights, biases, cell_list: Optional[bool] = None,
                  half: bool = False) -> torch.Tensor:
    """Both scales of a plain SA layer in one kernel (forward only).  radii / nsamples: 2 entries; weights / biases:
    6 BN-folded tensors (scale a's three layers, then scale b's).  Returns (B, cout_a + cout_b, M), the scales
    concatenated along the channel axis.  half: MLP contractions as fp16 single-pass products (pdab_sa_fused_pair_h: the
    TF32 class of the reference's cuDNN convolutions; same neighbour lists) instead of the fp32-level split products.  cell_list: bucket the points into a hashed cell list first so that a centre
    tests 27 cells instead of the whole cloud (same neighbour lists, bit-identical output); None = for clouds of >= 4096
    points (below that building the list costs more than the scan)."""
    assert xyz.is_cuda and xyz.is_contiguous() and new_xyz.is_contiguous()
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = 0 if features is None else features.shape[1]
    if features is not None:
        assert features.is_contiguous()
    dims_a = [3 + C] + [int(w.shape[0]) for w in weights[:3]]
    dims_b = [3 + C] + [int(w.shape[0]) for w in weights[3:]]
    for w, bvec in zip(weights, biases):
        assert w.is_cuda and w.is_contiguous() and w.dtype == torch.float32
        assert bvec.is_cuda and bvec.is_contiguous() and bvec.dtype == torch.float32
    out = torch.empty(B, dims_a[-1] + dims_b[-1], M, dtype=torch.float32, device=xyz.device)
    da, db = (ctypes.c_int * 4)(*dims_a), (ctypes.c_int * 4)(*dims_b)
    w_a = (ctypes.c_void_p * 6)(*[w.data_ptr() for w in weights])
    b_a = (ctypes.c_void_p * 6)(*[x.data_ptr() for x in biases])
    ws = None
    if cell_list or (cell_list is None and N >= 4096):
        ws = torch.empty(_lib.lib().pdab_sa_grid_workspace_bytes(B, N), dtype=torch.uint8, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.call("pdab_sa_fused_pair_h" if half else "pdab_sa_fused_pair", B, C, N, M, float(radii[0]), int(nsamples[0]), float(radii[1]), int(nsamples[1]),
                  xyz.data_ptr(), new_xyz.data_ptr(), features.data_ptr() if features is not None else None, da, db,
                  w_a, b_a, out.data_ptr(), ws.data_ptr() if ws is not None else None,
                  torch.cuda.current_stream(xyz.device).cuda_stream)
    return out


# --------------------------------------------------------------------------- grouper modules

class QueryAndGroup(nn.Module):
    """Ball query + grouping with centred xyz.  PB/pointnet2_utils.py:671-704."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_xyz


class QueryDilatedAndGroup(nn.Module):
    """PB/pointnet2_utils.py:706-741 (argument order radius_in, radius_out kept as in the reference,
    which forwards them to ball_query_dilated(max_radius, min_radius, ...) in that order)."""

    def __init__(self, radius_in: float, radius_out: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius_in, self.radius_out, self.nsample, self.use_xyz = radius_in, radius_out, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query_dilated(self.radius_in, self.radius_out, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_xyz


class QueryAndGroup_alone_grouped_density_directional(nn.Module):
    """PDA grouper: [xyz (not centred), Gaussian density, direction, features].
    PB/pointnet2_utils.py:557-614.  Inference (no grad) uses the fused kernel; with autograd
    enabled the differentiable composition of the elementary ops is used."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        if features is not None and self.use_xyz and not (torch.is_grad_enabled() and features.requires_grad):
            return pda_group(self.radius, self.nsample, xyz, new_xyz, features.contiguous())
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        centre = new_xyz.transpose(1, 2).unsqueeze(-1)
        offset = grouped_xyz - centre
        distances = torch.norm(offset, dim=1, keepdim=True)
        density = torch.exp(-distances ** 2 / (2 * self.radius ** 2)) / (2.5 * self.radius)
        direction = offset / self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            if self.use_xyz:
                return torch.cat([grouped_xyz, density, direction, grouped_features], dim=1)
            return grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_xyz


class GroupAll(nn.Module):
    """PB/pointnet2_utils.py:743-766."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz

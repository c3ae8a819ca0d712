"""Host side of the tcgen05 contractions (csrc/tc_gemm.cu, include/pdab.h "tensor-core contractions").

`PackedLinear` owns the packed shared-memory image of one weight matrix (built once per module in eval mode);
calling it runs  out = epilogue(x @ W^T + b)  on the 5th-generation tensor cores.  No fallback: CPU tensors are
rejected and a missing libpdab.so raises.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_STORE, EPI_RELU, EPI_ADD_LN, EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL, EPI_ATTN = range(6)


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


class PackedLinear:
    """y = epilogue(x W^T + b).  npass = 3: error-compensated 3xTF32 (fp32-level products); npass = 2: split-bf16
    "bf16x3" (hi/lo bf16 operands, 3 MMAs at twice the TF32 rate, ~2^-16 relative product error, fp32 range);
    npass = 1: plain TF32 (2^-11)."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], npass: int = 3, bn: Optional[int] = None,
                 xyz_last: int = 0):
        if not weight.is_cuda:
            raise RuntimeError("PackedLinear needs CUDA weights (pdanet_b200 has no CPU path)")
        w = weight.detach().float().contiguous()
        self.nout, self.k = int(w.shape[0]), int(w.shape[1])
        if npass == 2 and self.k % 8 and not (xyz_last == 3 and (self.k - 3) % 8 == 0):
            npass = 3   # split-bf16 packs 8 inputs per 16-byte chunk; odd widths (K = 12 position MLP) use 3xTF32
                        # (the SA gather prologue builds its [features (C % 8 == 0), xyz (3)] rows itself: C + 3 is fine there)
        self.npass = npass
        self.bn = bn if bn is not None else (256 if self.nout > 128 else 128)
        self.xyz_last = xyz_last
        self.bias = None if bias is None else bias.detach().float().contiguous()
        n = _lib.lib().pdab_tc_packed_floats(self.nout, self.k, npass, self.bn)
        self.packed = torch.empty(n, dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device):
            _lib.call("pdab_tc_pack_weights", self.nout, self.k, npass, self.bn, xyz_last, w.data_ptr(),
                      self.packed.data_ptr(), _stream(w))

    def __call__(self, x: torch.Tensor, epilogue: int = EPI_STORE, residual: Optional[torch.Tensor] = None,
                 norm: Optional[torch.nn.LayerNorm] = None, nsample: int = 0, out: Optional[torch.Tensor] = None):
        if not x.is_cuda:
            raise RuntimeError("tc_linear needs CUDA tensors")
        assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] == self.k and x.dtype == torch.float32
        rows = x.shape[0]
        pooled = epilogue in (EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL)
        out_rows = rows // nsample if pooled else rows
        if epilogue == EPI_ATTN:   # in_proj + neighbourhood attention: weight rows head-major [q_h | k_h | v_h] (attn_in_proj)
            assert self.bn == 192 and self.nout % 192 == 0 and out is None
            out = torch.empty(rows, self.nout // 3, dtype=torch.float32, device=x.device)
        if out is None:
            out = torch.empty(out_rows, self.nout, dtype=torch.float32, device=x.device)
        assert out.stride(1) == 1 and out.shape[0] == out_rows
        if residual is not None:
            assert residual.stride(1) == 1 and residual.shape == (rows, self.nout)
        with torch.cuda.device(x.device):
            _lib.call("pdab_tc_linear", rows, self.k, self.nout, self.npass, self.bn, epilogue, x.data_ptr(),
                      x.stride(0), self.packed.data_ptr(), None if self.bias is None else self.bias.data_ptr(),
                      None if residual is None else residual.data_ptr(),
                      0 if residual is None else residual.stride(0),
                      None if norm is None else norm.weight.data_ptr(),
                      None if norm is None else norm.bias.data_ptr(), float(norm.eps) if norm is not None else 0.0,
                      nsample, out.data_ptr(), out.stride(0), _stream(x))
        return out

    def sa_gather(self, idx: torch.Tensor, features_t: Optional[torch.Tensor], xyz: torch.Tensor,
                  new_xyz: torch.Tensor):
        """relu(W . [features_t[idx], xyz[idx] - new_xyz] + b) per (centre, sample) row; see pdab_tc_sa_gather_linear."""
        assert self.xyz_last == 3 and self.bn == 256
        B, M, ns = idx.shape
        N = xyz.shape[1]
        C = 0 if features_t is None else features_t.shape[2]
        assert C + 3 == self.k and idx.is_contiguous() and xyz.is_contiguous() and new_xyz.is_contiguous()
        assert features_t is None or features_t.is_contiguous()
        out = torch.empty(B * M * ns, self.nout, dtype=torch.float32, device=xyz.device)
        with torch.cuda.device(xyz.device):
            _lib.call("pdab_tc_sa_gather_linear", B, C, N, M, ns, self.nout, self.npass, idx.data_ptr(),
                      None if features_t is None else features_t.data_ptr(), xyz.data_ptr(), new_xyz.data_ptr(),
                      self.packed.data_ptr(), None if self.bias is None else self.bias.data_ptr(), out.data_ptr(),
                      out.stride(0), _stream(xyz))
        return out


def attn_in_proj(in_proj_weight: torch.Tensor, in_proj_bias: torch.Tensor, heads: int, npass: int = 2) -> "PackedLinear":
    """in_proj of nn.MultiheadAttention packed for the fused attention epilogue (EPI_ATTN, head_dim 64): the rows of the
    (3E, E) weight are regrouped head by head, [q_h | k_h | v_h] = one 192-column accumulator chunk per head, so that
    `lin(y, EPI_ATTN, nsample=ns)` returns softmax(q k^T / sqrt(hd)) v per neighbourhood, heads concatenated (T, E)."""
    E = in_proj_weight.shape[1]
    hd = E // heads
    assert hd == 64 and in_proj_weight.shape[0] == 3 * E
    order = torch.cat([torch.arange(part * E + h * hd, part * E + (h + 1) * hd) for h in range(heads) for part in range(3)])
    order = order.to(in_proj_weight.device)
    return PackedLinear(in_proj_weight.detach()[order], in_proj_bias.detach()[order], npass=npass, bn=192)

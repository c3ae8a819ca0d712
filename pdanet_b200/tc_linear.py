"""Host side of the tcgen05 contractions (csrc/tc_gemm.cu, include/pdab.h "tensor-core contractions").

`PackedLinear` owns the packed shared-memory image of one weight matrix (built once per module in eval mode);
calling it runs  out = epilogue(x @ W^T + b)  on the 5th-generation tensor cores.  No fallback: CPU tensors are
rejected and a missing libpdab.so raises.

Product modes (`npass`):
  3  error-compensated 3xTF32 (fp32-level products, 1e-6)             3 MMAs per k-step, fp32 tensors between kernels
  2  split-bf16 "bf16x3" (hi/lo bf16 operands, ~1e-5)                 3 MMAs per k-step at twice the TF32 rate
  1  plain TF32 (2^-11: the class of the reference's cuDNN convs)     1 MMA
  4  fp16 x fp16 single pass (2^-12 per operand, fp32 accumulation)   1 MMA at the bf16 rate; activations travel between
     kernels as fp16 (loaded by TMA), the transformer's residual streams as (hi, lo) fp16 plane pairs (`SplitHalf`,
     fp32-level) — see DESIGN.md §4 "precision": the residual stream, not the products, sets the block's output error.
"""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import _lib

EPI_STORE, EPI_RELU, EPI_ADD_LN, EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL, EPI_ATTN = range(6)
OUT_F32, OUT_F16, OUT_SPLIT = 0, 1, 2


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


class SplitHalf:
    """x = hi + lo as two fp16 (rows, cols) planes: 22 significand bits in the bytes of one fp32 matrix.  `hi` alone is
    the fp16 GEMM operand (read by TMA), hi + lo the residual stream."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor):
        assert hi.dtype == lo.dtype == torch.float16 and hi.shape == lo.shape
        self.hi, self.lo = hi, lo

    @classmethod
    def empty(cls, rows: int, cols: int, device) -> "SplitHalf":
        buf = torch.empty(2, rows, cols, dtype=torch.float16, device=device)
        return cls(buf[0], buf[1])

    @classmethod
    def from_float(cls, x: torch.Tensor) -> "SplitHalf":
        hi = x.half()
        return cls(hi, (x - hi.float()).half())

    def float(self) -> torch.Tensor:
        return self.hi.float() + self.lo.float()

    @property
    def shape(self):
        return self.hi.shape


class PackedLinear:
    """y = epilogue(x W^T + b); see the module docstring for `npass`."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], npass: int = 3, bn: Optional[int] = None,
                 xyz_last: int = 0):
        if not weight.is_cuda:
            raise RuntimeError("PackedLinear needs CUDA weights (pdanet_b200 has no CPU path)")
        w = weight.detach().float().contiguous()
        self.nout, self.k = int(w.shape[0]), int(w.shape[1])
        if npass in (2, 4) and self.k % 8 and not (xyz_last == 3 and (self.k - 3) % 8 == 0):
            npass = 3   # the 16-bit modes pack 8 inputs per 16-byte chunk; odd widths (K = 12 position MLP) use 3xTF32
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

    def __call__(self, x: torch.Tensor, epilogue: int = EPI_STORE,
                 residual: Union[torch.Tensor, SplitHalf, None] = None,
                 norm: Optional[torch.nn.LayerNorm] = None, nsample: int = 0, out: Optional[torch.Tensor] = None,
                 out_fmt: int = OUT_F32):
        """x (rows, k): fp32, or — npass = 4 only — fp16 (the TMA path).  `out_fmt` (npass = 4): OUT_F32, OUT_F16 or, for
        EPI_ADD_LN, OUT_SPLIT (returns a SplitHalf).  `residual`: fp32 tensor, or a SplitHalf with npass = 4."""
        if not x.is_cuda:
            raise RuntimeError("tc_linear needs CUDA tensors")
        if self.npass == 4:
            return self._call_h(x, epilogue, residual, norm, nsample, out, out_fmt)
        assert out_fmt == OUT_F32 and not isinstance(residual, SplitHalf), "16-bit tensors need npass = 4"
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

    def _call_h(self, x, epilogue, residual, norm, nsample, out, out_fmt):
        assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] == self.k and x.dtype in (torch.float32, torch.float16)
        rows = x.shape[0]
        a16 = x.dtype == torch.float16
        pooled = epilogue in (EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL)
        out_rows = rows // nsample if pooled else rows
        cols = self.nout // 3 if epilogue == EPI_ATTN else self.nout
        if epilogue == EPI_ATTN:
            assert self.bn == 192 and self.nout % 192 == 0
        if pooled:
            assert out_fmt == OUT_F32
        out_lo = None
        if out is None:
            if out_fmt == OUT_SPLIT:
                assert epilogue == EPI_ADD_LN
                split = SplitHalf.empty(out_rows, cols, x.device)
                out, out_lo = split.hi, split.lo
            else:
                out = torch.empty(out_rows, cols, dtype=torch.float16 if out_fmt == OUT_F16 else torch.float32,
                                  device=x.device)
        else:
            assert out_fmt != OUT_SPLIT and out.dtype == (torch.float16 if out_fmt == OUT_F16 else torch.float32)
        assert out.stride(1) == 1 and out.shape[0] == out_rows
        r_hi = r_lo = None
        ldr = 0
        if residual is not None:
            if isinstance(residual, SplitHalf):
                r_hi, r_lo = residual.hi, residual.lo
                assert r_hi.stride() == r_lo.stride()
            else:
                assert residual.dtype == torch.float32
                r_hi = residual
            assert r_hi.stride(1) == 1 and tuple(r_hi.shape) == (rows, self.nout)
            ldr = r_hi.stride(0)
        with torch.cuda.device(x.device):
            _lib.call("pdab_tc_linear_h", rows, self.k, self.nout, self.bn, epilogue, x.data_ptr(), x.stride(0), int(a16),
                      self.packed.data_ptr(), None if self.bias is None else self.bias.data_ptr(),
                      None if r_hi is None else r_hi.data_ptr(), None if r_lo is None else r_lo.data_ptr(), ldr,
                      None if norm is None else norm.weight.data_ptr(),
                      None if norm is None else norm.bias.data_ptr(), float(norm.eps) if norm is not None else 0.0,
                      nsample, out.data_ptr(), None if out_lo is None else out_lo.data_ptr(), out.stride(0), out_fmt,
                      _stream(x))
        return SplitHalf(out, out_lo) if out_lo is not None else out

    def sa_gather(self, idx: torch.Tensor, features_t: Optional[torch.Tensor], xyz: torch.Tensor,
                  new_xyz: torch.Tensor, out_fmt: int = OUT_F32):
        """relu(W . [features_t[idx], xyz[idx] - new_xyz] + b) per (centre, sample) row; see pdab_tc_sa_gather_linear."""
        assert self.xyz_last == 3 and self.bn == 256
        B, M, ns = idx.shape
        N = xyz.shape[1]
        C = 0 if features_t is None else features_t.shape[2]
        assert C + 3 == self.k and idx.is_contiguous() and xyz.is_contiguous() and new_xyz.is_contiguous()
        assert features_t is None or features_t.is_contiguous()
        assert out_fmt == OUT_F32 or (out_fmt == OUT_F16 and self.npass == 4)
        out = torch.empty(B * M * ns, self.nout, dtype=torch.float16 if out_fmt == OUT_F16 else torch.float32,
                          device=xyz.device)
        with torch.cuda.device(xyz.device):
            if self.npass == 4:
                _lib.call("pdab_tc_sa_gather_linear_h", B, C, N, M, ns, self.nout, idx.data_ptr(),
                          None if features_t is None else features_t.data_ptr(), xyz.data_ptr(), new_xyz.data_ptr(),
                          self.packed.data_ptr(), None if self.bias is None else self.bias.data_ptr(), out.data_ptr(),
                          out.stride(0), int(out_fmt == OUT_F16), _stream(xyz))
            else:
                _lib.call("pdab_tc_sa_gather_linear", B, C, N, M, ns, self.nout, self.npass, idx.data_ptr(),
                          None if features_t is None else features_t.data_ptr(), xyz.data_ptr(), new_xyz.data_ptr(),
                          self.packed.data_ptr(), None if self.bias is None else self.bias.data_ptr(), out.data_ptr(),
                          out.stride(0), _stream(xyz))
        return out


def ffn_fused_supported(e: int, nsample: int, *layers: "PackedLinear") -> bool:
    """Shapes `ffn_fused` covers (csrc/tc_ffn.cu): d_model 256, nsample 16 / 32, every layer in the fp16 single-pass mode."""
    return e == 256 and nsample in (16, 32) and all(l.npass == 4 for l in layers)


def ffn_fused(ctx: torch.Tensor, y: SplitHalf, out_proj: "PackedLinear", norm: torch.nn.LayerNorm, lin1: "PackedLinear",
              lin2: "PackedLinear", nsample: int) -> torch.Tensor:
    """max over each neighbourhood of (z + lin2(relu(lin1(z)))), z = norm(y + out_proj(ctx)), in ONE kernel
    (`pdab_tc_ffn_h`): ctx (T, 256) fp16, y (hi, lo) fp16 planes -> (T / nsample, 256) fp32; z and h never reach HBM."""
    T, E = ctx.shape
    assert ctx.dtype == torch.float16 and ctx.stride(1) == 1 and y.hi.shape == (T, E) and y.hi.stride() == y.lo.stride()
    assert (out_proj.nout, out_proj.k, out_proj.bn) == (E, E, 256) and (lin1.nout, lin1.k, lin1.bn) == (E // 2, E, 128)
    assert (lin2.nout, lin2.k, lin2.bn) == (E, E // 2, 256) and T % nsample == 0
    out = torch.empty(T // nsample, E, dtype=torch.float32, device=ctx.device)
    with torch.cuda.device(ctx.device):
        _lib.call("pdab_tc_ffn_h", T, E, nsample, ctx.data_ptr(), ctx.stride(0), y.hi.data_ptr(), y.lo.data_ptr(),
                  y.hi.stride(0), out_proj.packed.data_ptr(), out_proj.bias.data_ptr(), norm.weight.data_ptr(),
                  norm.bias.data_ptr(), float(norm.eps), lin1.packed.data_ptr(), lin1.bias.data_ptr(),
                  lin2.packed.data_ptr(), lin2.bias.data_ptr(), out.data_ptr(), out.stride(0), _stream(ctx))
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

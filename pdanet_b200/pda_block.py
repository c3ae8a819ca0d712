"""Eval-mode fast path of one PDA SA scale.

Same mathematics as `PointnetSAModuleMSG_WithSampling_Ellipsoid._scale` (the reference's
PB/pointnet2_modules.py:876-933), evaluated token-major so that nothing is permuted or re-laid-out between
the grouper kernel and the transformer:

  * the grouper kernel (`pdab_pda_group_tokens`) emits one contiguous row per (centre, neighbour) token;
  * every 1x1 convolution + eval BatchNorm becomes a folded linear layer on (tokens, channels) matrices;
  * every projection with K % 4 == 0 runs on the 5th-generation tensor cores through `PackedLinear`
    (csrc/tc_gemm.cu: tcgen05.mma, fp32 accumulators in TMEM) as an error-compensated split product —
    x = x_hi + x_lo, W = W_hi + W_lo split INSIDE the kernel, y = x_hi W_hi + x_hi W_lo + x_lo W_hi — with bf16 halves
    (`tc_passes` = 2, default: ~1e-5 relative, bf16 tensor rate) or TF32 halves (`tc_passes` = 3: fp32-level 1e-6),
    unlike plain TF32 (2^-11 = 5e-4, what the reference's cuDNN / cuBLAS calls use on tensor-core GPUs), which moves the
    PDA features by ~1e-3 and flips class-aware top-k picks downstream;
  * what used to be separate launches between the GEMMs is fused into their epilogues:
    out_proj + residual + LayerNorm2, linear1 + ReLU, linear2 + residual + max-pool over the neighbourhood;
  * attention itself (ns x ns per head, ns = 16/32) is one kernel on warp-level mma.sync TF32 with 3x hi/lo
    compensation, fp32 softmax (`pdab_group_attention`, csrc/pda_attn.cu).

The module's parameters are used as they are (state_dict unchanged); folded / packed copies are cached per module
and dropped on `.train()`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .tc_linear import (EPI_ADD_LN, EPI_ADD_MAXPOOL, EPI_ATTN, EPI_RELU, EPI_STORE, OUT_F16, OUT_SPLIT, PackedLinear,
                        SplitHalf, attn_in_proj, ffn_fused, ffn_fused_supported)


class LinearExact:
    """Layers whose K is not a multiple of 4 or that see only a handful of rows: plain IEEE fp32 addmm."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        self.w_t = weight.detach().float().t().contiguous()
        self.bias = bias.detach().float().contiguous()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            return torch.addmm(self.bias, x, self.w_t)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev


def _fold(conv, bn):
    from .pointnet2_modules import fold_conv_bn
    return fold_conv_bn(conv, bn)


class PDAScalePlan:
    """Folded / packed parameters of scale `i` of a PDA SA module."""

    def __init__(self, mod, i: int, npass: int = 2):
        self.radius = mod.groupers[i].radius
        self.ns = mod.nsamples[i]
        pm, gm, fc = mod.position_mlp[i], mod.global_mlps[i], mod.fin_conv[i]
        # npass = 4 (fp16 single pass) is for the transformer, whose residual streams stay at fp32 level.  The small direct
        # layers around it (position MLP on the unfused path, the two fin_conv layers on B*M pooled rows: < 1 % of the
        # block's FLOPs) have no residual to lean on: their operands would be rounded to 11 bits with nothing to hide it
        # (measured 5e-4 on the scale's output), so they keep the split-bf16 products (1e-5).
        small = 2 if npass == 4 else npass
        self.position = [PackedLinear(*_fold(pm[0], pm[1]), npass=small), PackedLinear(*_fold(pm[3], pm[4]), npass=small)]
        self.global_ = [LinearExact(*_fold(gm[0], gm[1])), LinearExact(*_fold(gm[3], gm[4]))]
        dn = mod.point_density[i].densitynet
        self.density = [LinearExact(*_fold(c, b)) for c, b in zip(dn.mlp_convs, dn.mlp_bns)]
        self.fin = [PackedLinear(*_fold(fc[0], fc[1]), npass=small), PackedLinear(*_fold(fc[3], fc[4]), npass=small)]
        tf = mod.Local_pointformer[i]
        attn = tf.self_attn
        self.heads = attn.num_heads
        self.in_proj = PackedLinear(attn.in_proj_weight, attn.in_proj_bias, npass=npass)
        # head_dim 64: in_proj and the attention core are ONE kernel (attention in the GEMM epilogue, qkv never stored)
        self.fused_attention = attn.embed_dim // attn.num_heads == 64
        self.in_proj_attn = (attn_in_proj(attn.in_proj_weight, attn.in_proj_bias, attn.num_heads, npass=npass)
                             if self.fused_attention else None)
        self.out_proj = PackedLinear(attn.out_proj.weight, attn.out_proj.bias, npass=npass)
        self.lin1 = PackedLinear(tf.linear1.weight, tf.linear1.bias, npass=npass)
        self.lin2 = PackedLinear(tf.linear2.weight, tf.linear2.bias, npass=npass)
        self.norm1, self.norm2 = tf.norm1, tf.norm2
        self.fused_ffn = True    # d_model 256, fp16 mode: the FFN half of the block as one kernel (csrc/tc_ffn.cu)
        # fused token encoder (csrc/pda_encode.cu): grouper + position MLP + DensityNet + token assembly + LayerNorm 1
        self.fused_encode = True
        self._enc_params = None
        self._enc_src = (_fold(pm[0], pm[1]), _fold(pm[3], pm[4]), [_fold(c, b) for c, b in zip(dn.mlp_convs, dn.mlp_bns)])

    @torch.no_grad()
    def __call__(self, ops, xyz, new_xyz, features_t, centre_feature_t, out=None):
        """xyz (B,N,3), new_xyz (B,M,3), features_t (B,N,C), centre_feature_t (B,M,C) -> (B, C_out, M).
        out: a (B*M, C_out) fp32 view (row stride free) the last layer writes its token-major result into — the module's
        token-major path hands in a column slice of the buffer that the aggregation layer reads (no cat, no transpose)."""
        self._out = out
        B, M, _ = new_xyz.shape
        ns = self.ns
        C = features_t.shape[2]
        G, T = B * M, B * M * ns
        if self.fused_encode and hasattr(ops, "pda_encode_ln") and ops.pda_encode_supported(C, ns):
            if self._enc_params is None:
                (w1, b1), (w2, b2), dens_layers = self._enc_src
                self._enc_params = ops.pda_encode_params(w1, b1, w2, b2, dens_layers, self.norm1.weight, self.norm1.bias)
            glob = torch.cat([new_xyz.reshape(G, 3), centre_feature_t.reshape(G, C)], dim=1)
            glob = F.relu_(self.global_[1](F.relu_(self.global_[0](glob))))
            if self.in_proj.npass == 4:   # fp16 single-pass path: tokens leave the encoder as (hi, lo) fp16 planes
                y = ops.pda_encode_ln(self.radius, ns, xyz, new_xyz, features_t, glob, self._enc_params, self.norm1.eps,
                                      split_half=True)
                return self._transformer_h(ops, y, B, M, ns)
            y = ops.pda_encode_ln(self.radius, ns, xyz, new_xyz, features_t, glob, self._enc_params, self.norm1.eps)
            return self._transformer(ops, y, B, M, ns)

        X = ops.pda_group_tokens(self.radius, ns, xyz, new_xyz, features_t).view(T, 8 + C)
        nbr, dens, direction = X[:, 0:3], X[:, 3], X[:, 4:7]

        # relative point position encoding -> position MLP (PB/pointnet2_modules.py:903-915)
        ctr = new_xyz.reshape(G, 1, 3).expand(G, ns, 3).reshape(T, 3)
        rppe = torch.cat([ctr, nbr, ctr - nbr, direction], dim=1)
        pos = self.position[1](self.position[0](rppe, EPI_RELU), EPI_RELU)

        # density re-weighting (PointConvDensitySetAbstraction + DensityNet, :958-1006)
        dg = dens.view(G, ns)
        scale = (dg / dg.max(dim=1, keepdim=True)[0]).reshape(T, 1)
        for lin in self.density:
            scale = F.relu_(lin(scale))

        # per-centre global feature, broadcast over the neighbourhood (:871-887)
        glob = torch.cat([new_xyz.reshape(G, 3), centre_feature_t.reshape(G, C)], dim=1)
        glob = F.relu_(self.global_[1](F.relu_(self.global_[0](glob))))

        # token assembly + LayerNorm 1 in one pass (csrc/pda_elem.cu)
        y = ops.pda_assemble_ln(pos, X, scale.reshape(T), glob, ns, self.norm1)
        if self.in_proj.npass == 4:
            return self._transformer_h(ops, SplitHalf.from_float(y), B, M, ns)
        return self._transformer(ops, y, B, M, ns)

    def _transformer(self, ops, y, B, M, ns):
        # pre-norm transformer over each neighbourhood (PB/PointFormer.py:28-38); residuals follow the LayerNorms
        if self.fused_attention and self.in_proj_attn.npass == self.in_proj.npass:
            ctx = self.in_proj_attn(y, EPI_ATTN, nsample=ns)                    # (T, E); qkv never exists
        else:
            qkv = self.in_proj(y, EPI_STORE)                                    # (T, 3E)
            ctx = ops.group_attention(qkv, ns, self.heads, npass=self.in_proj.npass)   # (T, E)
        z = self.out_proj(ctx, EPI_ADD_LN, residual=y, norm=self.norm2)         # LN2(y + attn)
        h = self.lin1(z, EPI_RELU)
        pooled = self.lin2(h, EPI_ADD_MAXPOOL, residual=z, nsample=ns)          # max_s (z + ffn), (:931)
        out = self.fin[1](self.fin[0](pooled, EPI_RELU), EPI_RELU, out=self._out)   # (G, C_out)
        return out if self._out is not None else out.view(B, M, -1).permute(0, 2, 1)

    def _transformer_h(self, ops, y: SplitHalf, B, M, ns):
        """The same block in the fp16 single-pass mode (tc_linear npass = 4): every GEMM operand is an fp16 matrix loaded
        by TMA, one tcgen05 MMA per k-step; the two residual streams (y, z) are (hi, lo) fp16 plane pairs, i.e. keep
        fp32-level precision — they, not the products, set the block's output error (tools/precision_study.py)."""
        if self.fused_attention:
            ctx = self.in_proj_attn(y.hi, EPI_ATTN, nsample=ns, out_fmt=OUT_F16)          # (T, E) fp16; qkv never exists
        else:
            qkv = self.in_proj(y.hi, EPI_STORE, out_fmt=OUT_F16)                         # (T, 3E) fp16
            ctx = ops.group_attention_h(qkv, ns, self.heads)                             # (T, E) fp16
        if self.fused_ffn and ffn_fused_supported(ctx.shape[1], ns, self.out_proj, self.lin1, self.lin2):
            # d_model 256: out_proj + LN2 + linear1 + ReLU + linear2 + residual + max-pool in ONE kernel, z and h on chip
            pooled = ffn_fused(ctx, y, self.out_proj, self.norm2, self.lin1, self.lin2, ns)
        else:
            z = self.out_proj(ctx, EPI_ADD_LN, residual=y, norm=self.norm2, out_fmt=OUT_SPLIT)   # LN2(y + attn) as (hi, lo)
            h = self.lin1(z.hi, EPI_RELU, out_fmt=OUT_F16)
            pooled = self.lin2(h, EPI_ADD_MAXPOOL, residual=z, nsample=ns)               # max_s (z + ffn), fp32 (G, E)
        out = self.fin[1](self.fin[0](pooled, EPI_RELU), EPI_RELU, out=self._out)        # (G, C_out)
        return out if self._out is not None else out.view(B, M, -1).permute(0, 2, 1)

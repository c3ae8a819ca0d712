"""Eval-mode fast path of one PDA SA scale (first step).

Same mathematics as `PointnetSAModuleMSG_WithSampling_Ellipsoid._scale` (the reference's
PB/pointnet2_modules.py:876-933), evaluated token-major so that nothing is permuted or re-laid-out between
the grouper kernel and the transformer:

  * the grouper kernel (`pdab_pda_group_tokens`) emits one contiguous row per (centre, neighbour) token;
  * every 1x1 convolution + eval BatchNorm becomes a folded `addmm` on (tokens, channels) matrices;
  * the four large projections of the pre-norm transformer (in_proj, out_proj, linear1, linear2) run on
    the tensor cores as error-compensated 3xTF32 products: x = x_hi + x_lo, W = W_hi + W_lo with the hi parts
    exactly representable in TF32, y = x_hi W_hi + x_hi W_lo + x_lo W_hi accumulated in fp32.  The dropped
    x_lo W_lo term and the truncation of the lo parts are O(2^-21) relative, i.e. fp32-level, unlike plain
    TF32 (2^-11) which moves the PDA features by ~1e-3 and flips class-aware top-k picks downstream;
  * attention itself (ns x ns per head, ns = 16/32) stays in IEEE fp32.

The module's parameters are used as they are (state_dict unchanged); folded / split copies are cached per module
and dropped on `.train()`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _split_tf32(x: torch.Tensor):
    """x = hi + lo with hi carrying the top 19 bits (exact in TF32) and lo the remainder (exact in fp32)."""
    hi = (x.view(torch.int32) & -8192).view(torch.float32)
    return hi, x - hi


class _TF32:
    """Scoped switch of torch's fp32-matmul mode (the flag is read at dispatch time)."""

    def __init__(self, on: bool):
        self.on = on

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.on

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class Linear3x:
    """y = x W^T + b with fp32-level accuracy on TF32 tensor cores (3 GEMMs)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        w_t = weight.detach().float().t().contiguous()  # (in, out)
        self.w_hi, self.w_lo = _split_tf32(w_t)
        self.w_hi, self.w_lo = self.w_hi.contiguous(), self.w_lo.contiguous()
        self.bias = bias.detach().float().contiguous()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.split_call(*_split_tf32(x))

    def split_call(self, x_hi: torch.Tensor, x_lo: torch.Tensor) -> torch.Tensor:
        with _TF32(True):
            y = torch.addmm(self.bias, x_hi, self.w_hi)
            y.addmm_(x_hi, self.w_lo)
            y.addmm_(x_lo, self.w_hi)
        return y


class LinearExact:
    """Small layers: plain IEEE fp32 addmm (K <= 16 or a handful of rows — not worth three GEMMs)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        self.w_t = weight.detach().float().t().contiguous()
        self.bias = bias.detach().float().contiguous()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        with _TF32(False):
            return torch.addmm(self.bias, x, self.w_t)


def _fold(conv, bn):
    from .pointnet2_modules import fold_conv_bn
    return fold_conv_bn(conv, bn)


def _mlp2(seq, big: bool):
    """(conv, bn, relu, conv, bn, relu) -> two folded linears."""
    mk = Linear3x if big else LinearExact
    return [mk(*_fold(seq[0], seq[1])), mk(*_fold(seq[3], seq[4]))]


class PDAScalePlan:
    """Folded / split parameters of scale `i` of a PDA SA module."""

    def __init__(self, mod, i: int):
        self.radius = mod.groupers[i].radius
        self.ns = mod.nsamples[i]
        self.position = _mlp2(mod.position_mlp[i], big=False)
        self.global_ = _mlp2(mod.global_mlps[i], big=False)
        dn = mod.point_density[i].densitynet
        self.density = [LinearExact(*_fold(c, b)) for c, b in zip(dn.mlp_convs, dn.mlp_bns)]
        self.fin = _mlp2(mod.fin_conv[i], big=True)
        tf = mod.Local_pointformer[i]
        attn = tf.self_attn
        self.heads = attn.num_heads
        self.in_proj = Linear3x(attn.in_proj_weight, attn.in_proj_bias)
        self.out_proj = Linear3x(attn.out_proj.weight, attn.out_proj.bias)
        self.lin1 = Linear3x(tf.linear1.weight, tf.linear1.bias)
        self.lin2 = Linear3x(tf.linear2.weight, tf.linear2.bias)
        self.norm1, self.norm2 = tf.norm1, tf.norm2

    @torch.no_grad()
    def __call__(self, ops, xyz, new_xyz, features_t, centre_feature_t):
        """xyz (B,N,3), new_xyz (B,M,3), features_t (B,N,C), centre_feature_t (B,M,C) -> (B, C_out, M)."""
        B, M, _ = new_xyz.shape
        ns = self.ns
        C = features_t.shape[2]
        G, T = B * M, B * M * ns
        X = ops.pda_group_tokens(self.radius, ns, xyz, new_xyz, features_t).view(T, 8 + C)
        nbr, dens, direction, feat = X[:, 0:3], X[:, 3], X[:, 4:7], X[:, 8:]

        # relative point position encoding -> position MLP (PB/pointnet2_modules.py:903-915)
        ctr = new_xyz.reshape(G, 1, 3).expand(G, ns, 3).reshape(T, 3)
        rppe = torch.cat([ctr, nbr, ctr - nbr, direction], dim=1)
        pos = F.relu_(self.position[1](F.relu_(self.position[0](rppe))))

        # density re-weighting (PointConvDensitySetAbstraction + DensityNet, :958-1006)
        dg = dens.view(G, ns)
        scale = (dg / dg.max(dim=1, keepdim=True)[0]).reshape(T, 1)
        for lin in self.density:
            scale = F.relu_(lin(scale))

        # per-centre global feature, broadcast over the neighbourhood (:871-887)
        glob = torch.cat([new_xyz.reshape(G, 3), centre_feature_t.reshape(G, C)], dim=1)
        glob = F.relu_(self.global_[1](F.relu_(self.global_[0](glob))))

        # token assembly + LayerNorm 1 + hi/lo split in one pass (csrc/pda_elem.cu); the normalised tokens exist
        # only as (hi, lo), which add back to the fp32 value exactly
        E, H = 4 * C, self.heads
        hd = E // H
        y_hi, y_lo = ops.pda_assemble_ln_split(pos, X, scale.reshape(T), glob, ns, self.norm1)

        # pre-norm transformer over each neighbourhood (PB/PointFormer.py:28-38); residuals follow the LayerNorms
        qkv = self.in_proj.split_call(y_hi, y_lo).view(G, ns, 3, H, hd)
        q = qkv[:, :, 0].permute(0, 2, 1, 3) * (1.0 / math.sqrt(hd))
        k = qkv[:, :, 1].permute(0, 2, 3, 1)
        v = qkv[:, :, 2].permute(0, 2, 1, 3)
        with _TF32(False):
            att = torch.softmax(torch.matmul(q, k), dim=-1)
            ctx = torch.matmul(att, v)                       # (G, H, ns, hd)
        ctx = ctx.permute(0, 2, 1, 3).reshape(T, E)
        z_hi, z_lo = ops.add_ln_split(y_hi, y_lo, self.out_proj(ctx), self.norm2)   # LN2(y + attn)
        h_hi, h_lo = ops.relu_split(self.lin1.split_call(z_hi, z_lo))
        pooled = ops.add_maxpool(z_hi, z_lo, self.lin2.split_call(h_hi, h_lo), G, ns)  # max_s (z + ffn), (:931)
        out = F.relu_(self.fin[1](F.relu_(self.fin[0](pooled))))  # (G, C_out)
        return out.view(B, M, -1).permute(0, 2, 1)

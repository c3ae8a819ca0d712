"""tcgen05 contraction kernels (csrc/tc_gemm.cu) against float64 references of the same op.

Tolerances (relative to the largest |reference| entry of the output):
  npass = 3 (error-compensated 3xTF32): 5e-6  — fp32-level, what the reference's nn.Linear computes;
  npass = 2 (split-bf16, "bf16x3")    : 4e-5  — 16-bit-mantissa products (hi/lo bf16 operands, 3 MMAs at the bf16 rate);
  npass = 1 (TF32 operands)           : 2e-3  — TF32 level, what the reference's cuDNN 1x1 convolutions compute.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {3: 5e-6, 2: 4e-5, 1: 2e-3}


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _rel(out, ref):
    ref = ref.double()
    return float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _mk(rows, k, nout, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, k, generator=g)
    w = torch.randn(nout, k, generator=g) / k ** 0.5
    b = torch.randn(nout, generator=g)
    return x, w, b


@pytest.mark.parametrize("npass", [3, 2, 1])
@pytest.mark.parametrize("rows,k,nout", [(128, 32, 128), (1000, 256, 768), (4096, 72, 256), (333, 512, 1536),
                                         (2048, 128, 64), (5000, 264, 512)])
def test_linear_store_and_relu(npass, rows, k, nout):
    from pdanet_b200.tc_linear import PackedLinear, EPI_STORE, EPI_RELU
    dev = _dev()
    x, w, b = _mk(rows, k, nout, seed=rows + k)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=npass)
    ref = x.double() @ w.double().t() + b.double()
    out = lin(x.to(dev), EPI_STORE).cpu()
    assert _rel(out, ref) < TOL[npass]
    out = lin(x.to(dev), EPI_RELU).cpu()
    assert _rel(out, ref.clamp_min(0)) < TOL[npass]
    # no bias, strided input rows
    lin2 = PackedLinear(w.to(dev), None, npass=npass)
    xp = torch.zeros(rows, k + 8)
    xp[:, :k] = x
    out = lin2(xp.to(dev)[:, :k], EPI_STORE).cpu()
    assert _rel(out, x.double() @ w.double().t()) < TOL[npass]


@pytest.mark.parametrize("npass", [3, 2, 1])
@pytest.mark.parametrize("e", [256, 512])
def test_linear_add_layernorm(npass, e):
    from pdanet_b200.tc_linear import PackedLinear, EPI_ADD_LN
    dev = _dev()
    rows = 1500
    x, w, b = _mk(rows, e, e, seed=e)
    res = torch.randn(rows, e, generator=torch.Generator().manual_seed(5))
    norm = torch.nn.LayerNorm(e)
    with torch.no_grad():
        norm.weight.copy_(torch.rand(e) + 0.5)
        norm.bias.copy_(torch.randn(e) * 0.1)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=npass)
    ref = torch.nn.functional.layer_norm(x.double() @ w.double().t() + b.double() + res.double(), (e,),
                                         norm.weight.double(), norm.bias.double(), norm.eps)
    out = lin(x.to(dev), EPI_ADD_LN, residual=res.to(dev), norm=norm.to(dev)).cpu()
    assert _rel(out, ref) < TOL[npass] * 2


@pytest.mark.parametrize("npass", [3, 2, 1])
@pytest.mark.parametrize("ns", [16, 32])
def test_linear_maxpool_epilogues(npass, ns):
    from pdanet_b200.tc_linear import PackedLinear, EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL
    dev = _dev()
    groups, k, nout = 100, 128, 512
    rows = groups * ns
    x, w, b = _mk(rows, k, nout, seed=ns)
    res = torch.randn(rows, nout, generator=torch.Generator().manual_seed(6))
    lin = PackedLinear(w.to(dev), b.to(dev), npass=npass)
    y = x.double() @ w.double().t() + b.double()
    ref = (y + res.double()).view(groups, ns, nout).max(dim=1)[0]
    out = lin(x.to(dev), EPI_ADD_MAXPOOL, residual=res.to(dev), nsample=ns).cpu()
    assert out.shape == (groups, nout)
    assert _rel(out, ref) < TOL[npass]
    ref = y.clamp_min(0).view(groups, ns, nout).max(dim=1)[0]
    out = lin(x.to(dev), EPI_RELU_MAXPOOL, nsample=ns).cpu()
    assert _rel(out, ref) < TOL[npass]


@pytest.mark.parametrize("npass", [3, 2, 1])
def test_sa_gather_linear(npass):
    """Gather prologue == grouping_operation x2 + centre subtraction + cat + first 1x1 conv (reference channel order)."""
    from pdanet_b200.tc_linear import PackedLinear
    dev = _dev()
    B, N, M, ns, C, nout = 3, 512, 200, 16, 64, 256
    g = torch.Generator().manual_seed(11)
    xyz = torch.rand(B, N, 3, generator=g) * 10
    new_xyz = torch.rand(B, M, 3, generator=g) * 10
    feat = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, M, ns), generator=g, dtype=torch.int32)
    w = torch.randn(nout, 3 + C, generator=g) / (3 + C) ** 0.5   # reference input order: [xyz(3), features(C)]
    b = torch.randn(nout, generator=g)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=npass, bn=256, xyz_last=3)
    out = lin.sa_gather(idx.to(dev), feat.transpose(1, 2).contiguous().to(dev), xyz.to(dev), new_xyz.to(dev)).cpu()
    li = idx.long()
    gx = torch.gather(xyz.unsqueeze(1).expand(B, M, N, 3), 2, li.unsqueeze(-1).expand(B, M, ns, 3)) - new_xyz.unsqueeze(2)
    gf = torch.gather(feat.transpose(1, 2).unsqueeze(1).expand(B, M, N, C), 2, li.unsqueeze(-1).expand(B, M, ns, C))
    rows = torch.cat([gx, gf], dim=-1).reshape(B * M * ns, 3 + C).double()
    ref = (rows @ w.double().t() + b.double()).clamp_min(0)
    assert _rel(out, ref) < TOL[npass]


def test_large_persistent_many_items():
    """More work items than SMs: the persistent loop, both accumulator stages and the smem ring wrap many times."""
    from pdanet_b200.tc_linear import PackedLinear, EPI_STORE
    dev = _dev()
    rows, k, nout = 70000, 256, 768
    x, w, b = _mk(rows, k, nout, seed=3)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=3)
    out = lin(x.to(dev), EPI_STORE)
    ref = (x.to(dev).double() @ w.to(dev).double().t() + b.to(dev).double())
    assert _rel(out.cpu(), ref.cpu()) < TOL[3]


@pytest.mark.parametrize("npass", [3, 2])
@pytest.mark.parametrize("ns,hd", [(16, 64), (32, 64), (16, 128), (32, 128)])
def test_group_attention_vs_torch_mha(ns, hd, npass):
    """pdab_group_attention == the attention core of nn.MultiheadAttention on (ns, groups, E) sequences."""
    from pdanet_b200 import pointnet2_utils as ops
    dev = _dev()
    heads, groups = 4, 37
    E = heads * hd
    g = torch.Generator().manual_seed(ns + hd)
    qkv = torch.randn(groups * ns, 3 * E, generator=g) * 1.5
    ctx = ops.group_attention(qkv.to(dev), ns, heads, npass=npass).cpu()
    q, k, v = [t.double().view(groups, ns, heads, hd).permute(0, 2, 1, 3) for t in qkv.split(E, dim=1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(groups * ns, E)
    assert _rel(ctx, ref) < (5e-6 if npass == 3 else 2 * TOL[2])


@pytest.mark.parametrize("npass", [3, 2, 1])
@pytest.mark.parametrize("ns", [16, 32])
@pytest.mark.parametrize("pairs", [1, 0])
def test_in_proj_attention_epilogue(npass, ns, pairs):
    """EPI_ATTN (in_proj + neighbourhood attention in one kernel, head_dim 64) == nn.MultiheadAttention's in_proj followed
    by its attention core, in float64.  Ragged row count (last pair tile half empty), CTA pairs and single CTAs."""
    from pdanet_b200 import _lib
    from pdanet_b200.tc_linear import EPI_ATTN, attn_in_proj
    dev = _dev()
    heads, hd = 4, 64
    E = heads * hd
    groups = 4000 // ns * 3 + 5
    rows = groups * ns
    g = torch.Generator().manual_seed(ns + npass)
    x = torch.randn(rows, E, generator=g)
    w = torch.randn(3 * E, E, generator=g) / E ** 0.5 * 1.5
    b = torch.randn(3 * E, generator=g) * 0.3
    try:
        _lib.lib().pdab_set_cta_pairs(pairs)
        lin = attn_in_proj(w.to(dev), b.to(dev), heads, npass=npass)
        ctx = lin(x.to(dev), EPI_ATTN, nsample=ns).cpu()
    finally:
        _lib.lib().pdab_set_cta_pairs(1)
    qkv = x.double() @ w.double().t() + b.double()
    q, k, v = [t.view(groups, ns, heads, hd).permute(0, 2, 1, 3) for t in qkv.split(E, dim=1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(rows, E)
    assert ctx.shape == ref.shape
    assert _rel(ctx, ref) < 3 * TOL[npass]     # three chained contractions


@pytest.mark.parametrize("tc_passes", [3, 2, 1])
def test_wide_sa_scale_matches_unfused_module(tc_passes):
    """Plain SA layer with wide MLPs (the L5 shape): tensor-core path vs the reference statement order
    (QueryAndGroup -> Conv2d/BN/ReLU x3 -> max_pool2d) of the same module with the same parameters."""
    from pdanet_b200.pointnet2_modules import PointnetSAModuleMSG_WithSampling
    dev = _dev()
    torch.manual_seed(0)
    mod = PointnetSAModuleMSG_WithSampling(
        npoint_list=[64], sample_range_list=[-1], sample_type_list=["D-FPS"], radii=[4.8, 6.4], nsamples=[16, 32],
        mlps=[[64, 64, 64, 128], [64, 64, 128, 256]], aggregation_mlp=[128], confidence_mlp=[], num_class=3).to(dev).eval()
    with torch.no_grad():
        for m in mod.modules():  # non-trivial BN statistics so that the folding is exercised
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    mod.tc_passes = tc_passes
    g = torch.Generator().manual_seed(1)
    B, N = 2, 512
    xyz = (torch.rand(B, N, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])).to(dev)
    feats = torch.randn(B, 64, N, generator=g).to(dev)
    ctr = (torch.rand(B, 100, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])).to(dev)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            _, fused, _, _ = mod(xyz, feats, None, ctr_xyz=ctr)
            assert len(mod._wide) == 2, "tensor-core path was not taken"
            mod.ops = type("NoFused", (), {k: getattr(mod.ops, k) for k in ("ball_query", "grouping_operation",
                           "gather_operation", "furthest_point_sample", "QueryAndGroup")})
            _, plain, _, _ = mod(xyz, feats, None, ctr_xyz=ctr)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    err = float((fused - plain).abs().max() / plain.abs().max())
    assert err < {3: 2e-5, 2: 2e-4, 1: 3e-3}[tc_passes], err


@pytest.mark.parametrize("npass", [3, 2, 1])
def test_cta_pairs_equal_single_ctas(npass):
    """cta_group::2 CTA pairs (M = 256 across two SMs) are a schedule, not arithmetic: every epilogue / prologue gives
    bit-identical results with pairs on and off (same MMAs in the same order per output element)."""
    from pdanet_b200 import _lib
    from pdanet_b200.tc_linear import (PackedLinear, EPI_STORE, EPI_RELU, EPI_ADD_LN, EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL)
    dev = _dev()
    rows, k = 6000 * 2 + 96, 256          # not a multiple of 256: the last pair tile is ragged, one CTA of it empty
    x, w, b = _mk(rows, k, 512, seed=9)
    res = torch.randn(rows, 512, generator=torch.Generator().manual_seed(10))
    norm = torch.nn.LayerNorm(512).to(dev)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=npass)
    lin128 = PackedLinear(w[:128].to(dev), b[:128].to(dev), npass=npass)
    xd, rd = x.to(dev), res.to(dev)
    outs = {}
    try:
        for pairs in (1, 0):
            _lib.lib().pdab_set_cta_pairs(pairs)
            outs[pairs] = [
                lin(xd, EPI_STORE), lin(xd, EPI_RELU), lin(xd, EPI_ADD_LN, residual=rd, norm=norm),
                lin(xd, EPI_ADD_MAXPOOL, residual=rd, nsample=32), lin(xd, EPI_RELU_MAXPOOL, nsample=16),
                lin128(xd, EPI_RELU),
            ]
    finally:
        _lib.lib().pdab_set_cta_pairs(1)
    for a, c in zip(outs[1], outs[0]):
        assert torch.equal(a, c)
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(outs[1][0].cpu(), ref) < TOL[npass]


# ------------------------------------------------------------------------------------------------------------------
# npass = 4: fp16 x fp16 single pass, fp16 activations between kernels (TMA-fed A operand), (hi, lo) fp16 residuals.
# Kernel correctness is checked against float64 references built from the SAME fp16-rounded operands (what the tensor
# core multiplies): the only differences left are fp32 accumulation order and, for fp16 outputs, the final rounding
# (2^-11 of each element).  The distance to the un-rounded fp32 math — the precision class of the mode — is checked
# separately at 1e-3 (north_star's feature bar).

H_ACC = 1e-5          # fp32 accumulation of exact fp16 x fp16 products
H_OUT16 = 6e-4        # + one fp16 rounding of the output


def _h(t):            # what the kernel sees of an fp32 tensor in this mode
    return t.half().double()


@pytest.mark.parametrize("rows,k,nout", [(128, 64, 128), (1000, 256, 768), (4096, 72, 256), (333, 512, 1536),
                                         (2048, 128, 64), (5000, 264, 512), (70000, 256, 256)])
def test_h16_linear_tma_and_rows(rows, k, nout):
    from pdanet_b200.tc_linear import PackedLinear, EPI_STORE, EPI_RELU, OUT_F16
    dev = _dev()
    x, w, b = _mk(rows, k, nout, seed=rows + k)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=4)
    assert lin.npass == 4
    ref = _h(x) @ _h(w).t() + b.double()
    x16 = x.to(dev).half()
    assert _rel(lin(x16, EPI_STORE).cpu(), ref) < H_ACC                        # TMA operand, fp32 out
    assert _rel(lin(x16, EPI_RELU, out_fmt=OUT_F16).cpu(), ref.clamp_min(0)) < H_OUT16
    assert _rel(lin(x.to(dev), EPI_RELU).cpu(), ref.clamp_min(0)) < H_ACC      # fp32 rows converted by the producer warps
    assert _rel(lin(x16, EPI_STORE).cpu(), x.double() @ w.double().t() + b.double()) < 1e-3   # the mode's precision class
    # strided fp16 rows (lda > k), no bias
    lin2 = PackedLinear(w.to(dev), None, npass=4)
    xp = torch.zeros(rows, k + 24, dtype=torch.float16, device=dev)
    xp[:, :k] = x16
    assert _rel(lin2(xp[:, :k], EPI_STORE).cpu(), _h(x) @ _h(w).t()) < H_ACC


@pytest.mark.parametrize("e", [256, 512])
@pytest.mark.parametrize("split_res", [True, False])
def test_h16_add_layernorm(e, split_res):
    from pdanet_b200.tc_linear import PackedLinear, SplitHalf, EPI_ADD_LN, OUT_SPLIT, OUT_F16
    dev = _dev()
    rows = 1500
    x, w, b = _mk(rows, e, e, seed=e)
    res = torch.randn(rows, e, generator=torch.Generator().manual_seed(5))
    norm = torch.nn.LayerNorm(e)
    with torch.no_grad():
        norm.weight.copy_(torch.rand(e) + 0.5)
        norm.bias.copy_(torch.randn(e) * 0.1)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=4)
    r_dev = SplitHalf.from_float(res.to(dev)) if split_res else res.to(dev)
    r_ref = r_dev.float().cpu().double() if split_res else res.double()
    assert (r_ref - res.double()).abs().max() < 1e-6                            # (hi, lo) carries ~22 bits
    ref = torch.nn.functional.layer_norm(_h(x) @ _h(w).t() + b.double() + r_ref, (e,), norm.weight.double(),
                                         norm.bias.double(), norm.eps)
    z = lin(x.to(dev).half(), EPI_ADD_LN, residual=r_dev, norm=norm.to(dev), out_fmt=OUT_SPLIT)
    assert isinstance(z, SplitHalf) and z.hi.dtype == torch.float16
    assert _rel(z.float().cpu(), ref) < 2 * H_ACC
    out = lin(x.to(dev).half(), EPI_ADD_LN, residual=r_dev, norm=norm.to(dev)).cpu()          # fp32 out
    assert _rel(out, ref) < 2 * H_ACC
    out = lin(x.to(dev).half(), EPI_ADD_LN, residual=r_dev, norm=norm.to(dev), out_fmt=OUT_F16).cpu()
    assert _rel(out, ref) < H_OUT16


@pytest.mark.parametrize("ns", [16, 32, 64])
def test_h16_maxpool_epilogues(ns):
    from pdanet_b200.tc_linear import PackedLinear, SplitHalf, EPI_ADD_MAXPOOL, EPI_RELU_MAXPOOL
    dev = _dev()
    groups, k, nout = 300, 128, 512
    rows = groups * ns
    x, w, b = _mk(rows, k, nout, seed=ns)
    res = torch.randn(rows, nout, generator=torch.Generator().manual_seed(6))
    lin = PackedLinear(w.to(dev), b.to(dev), npass=4)
    y = _h(x) @ _h(w).t() + b.double()
    x16 = x.to(dev).half()
    ref = y.clamp_min(0).view(groups, ns, nout).max(dim=1)[0]
    out = lin(x16, EPI_RELU_MAXPOOL, nsample=ns).cpu()
    assert out.shape == (groups, nout) and out.dtype == torch.float32
    assert _rel(out, ref) < H_ACC
    if ns == 64:      # a 64-row neighbourhood spans two TMEM quadrants: only the ReLU form (atomicMax on bit patterns)
        return
    r = SplitHalf.from_float(res.to(dev))
    ref = (y + r.float().cpu().double()).view(groups, ns, nout).max(dim=1)[0]
    assert _rel(lin(x16, EPI_ADD_MAXPOOL, residual=r, nsample=ns).cpu(), ref) < H_ACC
    ref = (y + res.double()).view(groups, ns, nout).max(dim=1)[0]
    assert _rel(lin(x16, EPI_ADD_MAXPOOL, residual=res.to(dev), nsample=ns).cpu(), ref) < H_ACC


@pytest.mark.parametrize("B,M", [(3, 200), (9, 5)])   # tiles that cross one scene boundary / several (80 rows per scene)
def test_h16_sa_gather_linear(B, M):
    from pdanet_b200.tc_linear import PackedLinear, OUT_F16
    dev = _dev()
    N, ns, C, nout = 512, 16, 64, 256
    g = torch.Generator().manual_seed(11)
    xyz = torch.rand(B, N, 3, generator=g) * 10
    new_xyz = torch.rand(B, M, 3, generator=g) * 10
    feat = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, M, ns), generator=g, dtype=torch.int32)
    w = torch.randn(nout, 3 + C, generator=g) / (3 + C) ** 0.5
    b = torch.randn(nout, generator=g)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=4, bn=256, xyz_last=3)
    assert lin.npass == 4
    li = idx.long()
    gx = torch.gather(xyz.unsqueeze(1).expand(B, M, N, 3), 2, li.unsqueeze(-1).expand(B, M, ns, 3)) - new_xyz.unsqueeze(2)
    gf = torch.gather(feat.transpose(1, 2).unsqueeze(1).expand(B, M, N, C), 2, li.unsqueeze(-1).expand(B, M, ns, C))
    rows = torch.cat([gx, gf], dim=-1).reshape(B * M * ns, 3 + C)
    ref = (_h(rows) @ _h(w).t() + b.double()).clamp_min(0)
    args = (idx.to(dev), feat.transpose(1, 2).contiguous().to(dev), xyz.to(dev), new_xyz.to(dev))
    assert _rel(lin.sa_gather(*args).cpu(), ref) < H_ACC
    out16 = lin.sa_gather(*args, out_fmt=OUT_F16)
    assert out16.dtype == torch.float16 and _rel(out16.cpu(), ref) < H_OUT16


@pytest.mark.parametrize("ns", [16, 32])
@pytest.mark.parametrize("pairs", [1, 0])
def test_h16_in_proj_attention_epilogue(ns, pairs):
    """EPI_ATTN in the fp16 mode: TMA-fed in_proj, fp16 m16n8k16 attention contractions (one MMA per k-step), fp16 ctx."""
    from pdanet_b200 import _lib
    from pdanet_b200.tc_linear import EPI_ATTN, OUT_F16, attn_in_proj
    dev = _dev()
    heads, hd = 4, 64
    E = heads * hd
    groups = 4000 // ns * 3 + 5
    rows = groups * ns
    g = torch.Generator().manual_seed(ns + 4)
    x = torch.randn(rows, E, generator=g)
    w = torch.randn(3 * E, E, generator=g) / E ** 0.5 * 1.5
    b = torch.randn(3 * E, generator=g) * 0.3
    try:
        _lib.lib().pdab_set_cta_pairs(pairs)
        lin = attn_in_proj(w.to(dev), b.to(dev), heads, npass=4)
        ctx16 = lin(x.to(dev).half(), EPI_ATTN, nsample=ns, out_fmt=OUT_F16)
        ctx32 = lin(x.to(dev).half(), EPI_ATTN, nsample=ns)
    finally:
        _lib.lib().pdab_set_cta_pairs(1)
    qkv = _h(x) @ _h(w).t() + b.double()
    q, k, v = [t.view(groups, ns, heads, hd).permute(0, 2, 1, 3) for t in qkv.split(E, dim=1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(rows, E)
    assert ctx16.dtype == torch.float16 and ctx16.shape == ref.shape
    # q, k, v and the probabilities are rounded to fp16 for the attention MMAs: 2^-12 each on O(1) values
    assert _rel(ctx32.cpu(), ref) < 1.5e-3
    assert _rel(ctx16.cpu(), ref) < 2e-3
    assert float((ctx16.float() - ctx32).abs().max() / ctx32.abs().max()) < H_OUT16


def test_h16_cta_pairs_equal_single_ctas():
    from pdanet_b200 import _lib
    from pdanet_b200.tc_linear import (PackedLinear, SplitHalf, EPI_STORE, EPI_RELU, EPI_ADD_LN, EPI_ADD_MAXPOOL,
                                       EPI_RELU_MAXPOOL, OUT_F16, OUT_SPLIT)
    dev = _dev()
    rows, k = 6000 * 2 + 96, 256
    x, w, b = _mk(rows, k, 512, seed=9)
    res = SplitHalf.from_float(torch.randn(rows, 512, generator=torch.Generator().manual_seed(10)).to(dev))
    norm = torch.nn.LayerNorm(512).to(dev)
    lin = PackedLinear(w.to(dev), b.to(dev), npass=4)
    lin128 = PackedLinear(w[:128].to(dev), b[:128].to(dev), npass=4)
    xd = x.to(dev).half()
    outs = {}
    try:
        for pairs in (1, 0):
            _lib.lib().pdab_set_cta_pairs(pairs)
            z = lin(xd, EPI_ADD_LN, residual=res, norm=norm, out_fmt=OUT_SPLIT)
            outs[pairs] = [
                lin(xd, EPI_STORE), lin(xd, EPI_RELU, out_fmt=OUT_F16), z.hi, z.lo,
                lin(xd, EPI_ADD_MAXPOOL, residual=res, nsample=32), lin(xd, EPI_RELU_MAXPOOL, nsample=16),
                lin128(xd, EPI_RELU, out_fmt=OUT_F16),
            ]
    finally:
        _lib.lib().pdab_set_cta_pairs(1)
    for a, c in zip(outs[1], outs[0]):
        assert torch.equal(a, c)
    assert _rel(outs[1][0].cpu(), _h(x) @ _h(w).t() + b.double()) < H_ACC


@pytest.mark.parametrize("nsamples", [[16, 32], [16, 32, 64]])
def test_h16_wide_sa_scale_matches_unfused_module(nsamples):
    """The L5 shapes (KITTI: two scales; ONCE: three, the third with nsample 64) in the fp16 single-pass mode against the
    module's reference statement order in fp32: inside the 1e-3 feature bar, and no scale falls back to groupers[i]."""
    from pdanet_b200.pointnet2_modules import PointnetSAModuleMSG_WithSampling
    dev = _dev()
    torch.manual_seed(0)
    n = len(nsamples)
    mod = PointnetSAModuleMSG_WithSampling(
        npoint_list=[64], sample_range_list=[-1], sample_type_list=["D-FPS"], radii=[4.8, 6.4, 8.4][:n], nsamples=nsamples,
        mlps=[[64, 64, 64, 128], [64, 64, 128, 256], [64, 64, 128, 128]][:n], aggregation_mlp=[128], confidence_mlp=[],
        num_class=3).to(dev).eval()
    with torch.no_grad():
        for m in mod.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    mod.tc_passes = 4
    g = torch.Generator().manual_seed(1)
    B, N = 2, 512
    xyz = (torch.rand(B, N, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])).to(dev)
    feats = torch.randn(B, 64, N, generator=g).to(dev)
    ctr = (torch.rand(B, 100, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])).to(dev)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            _, fused, _, _ = mod(xyz, feats, None, ctr_xyz=ctr)
            assert len(mod._wide) == n, "a scale fell back to groupers[i] + eager convolutions"
            mod.ops = type("NoFused", (), {k: getattr(mod.ops, k) for k in ("ball_query", "grouping_operation",
                           "gather_operation", "furthest_point_sample", "QueryAndGroup")})
            _, plain, _, _ = mod(xyz, feats, None, ctr_xyz=ctr)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    err = float((fused - plain).abs().max() / plain.abs().max())
    assert err < 1e-3, err


@pytest.mark.parametrize("ns,hd", [(16, 64), (32, 64), (16, 128), (32, 128)])
def test_h16_group_attention_vs_torch_mha(ns, hd):
    """pdab_group_attention_h (fp16 qkv / ctx, fp16 MMAs, fp32 softmax) == the attention core of nn.MultiheadAttention on
    the fp16-rounded q, k, v in float64; the probabilities are rounded to fp16 for P V and ctx to fp16 on the way out."""
    from pdanet_b200 import pointnet2_utils as ops
    dev = _dev()
    heads, groups = 4, 37
    E = heads * hd
    g = torch.Generator().manual_seed(ns + hd)
    qkv = (torch.randn(groups * ns, 3 * E, generator=g) * 1.5).half()
    ctx = ops.group_attention_h(qkv.to(dev), ns, heads).cpu()
    assert ctx.dtype == torch.float16
    q, k, v = [t.double().view(groups, ns, heads, hd).permute(0, 2, 1, 3) for t in qkv.split(E, dim=1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(groups * ns, E)
    assert _rel(ctx, ref) < 1.5e-3


@pytest.mark.parametrize("ns", [16, 32])
@pytest.mark.parametrize("groups", [3, 700, 9001])
def test_h16_ffn_fused_equals_three_launch_chain(ns, groups):
    """pdab_tc_ffn_h (out_proj + LN2 + linear1 + ReLU + linear2 + residual + max-pool in one kernel, z / h on chip) against
    (a) the three-launch chain of pdab_tc_linear_h it replaces and (b) a float64 statement of the same mathematics on the
    fp16-rounded operands.  Row counts: less than one CTA tile, ragged last pair tile, many tiles per CTA."""
    from pdanet_b200.tc_linear import (PackedLinear, SplitHalf, EPI_ADD_LN, EPI_ADD_MAXPOOL, EPI_RELU, OUT_F16, OUT_SPLIT,
                                       ffn_fused, ffn_fused_supported)
    dev = _dev()
    E = 256
    T = groups * ns
    g = torch.Generator().manual_seed(groups + ns)
    ctx = (torch.randn(T, E, generator=g) * 0.7).to(dev).half()
    y32 = torch.randn(T, E, generator=g).to(dev)
    y = SplitHalf.from_float(y32)
    wo, bo = torch.randn(E, E, generator=g) / E ** 0.5, torch.randn(E, generator=g) * 0.1
    w1, b1 = torch.randn(E // 2, E, generator=g) / E ** 0.5, torch.randn(E // 2, generator=g) * 0.1
    w2, b2 = torch.randn(E, E // 2, generator=g) / (E // 2) ** 0.5, torch.randn(E, generator=g) * 0.1
    norm = torch.nn.LayerNorm(E).to(dev)
    with torch.no_grad():
        norm.weight.copy_(torch.rand(E, generator=g) + 0.5)
        norm.bias.copy_(torch.randn(E, generator=g) * 0.1)
    lo, l1, l2 = (PackedLinear(w.to(dev), b.to(dev), npass=4) for w, b in ((wo, bo), (w1, b1), (w2, b2)))
    assert ffn_fused_supported(E, ns, lo, l1, l2)
    fused = ffn_fused(ctx, y, lo, norm, l1, l2, ns)
    z = lo(ctx, EPI_ADD_LN, residual=y, norm=norm, out_fmt=OUT_SPLIT)
    h = l1(z.hi, EPI_RELU, out_fmt=OUT_F16)
    chain = l2(h, EPI_ADD_MAXPOOL, residual=z, nsample=ns)
    assert fused.shape == chain.shape == (groups, E)
    # Same products, same operand rounding; the two differ in the summation order of the LayerNorm statistics (row-split
    # against column-split epilogue warps), which moves z by an fp32 ulp and so, now and then, flips the fp16 rounding of a
    # z element that feeds linear1 (2^-11 of one of 256 inputs): measured 1e-5 typically, 1.4e-4 on the worst of 2 M outputs
    assert _rel(fused.cpu(), chain.cpu()) < 3e-4
    d = lambda t: t.detach().cpu().double()
    zz = torch.nn.functional.layer_norm(d(ctx) @ _h(wo).t() + bo.double() + d(y.float()), (E,), d(norm.weight), d(norm.bias), norm.eps)
    hh = torch.relu(_h(zz.float()) @ _h(w1).t() + b1.double())
    ref = (zz + _h(hh.float()) @ _h(w2).t() + b2.double()).view(groups, ns, E).max(dim=1)[0]
    assert _rel(fused.cpu(), ref) < 3e-4

#!/usr/bin/env python
"""Run one op a few times (target for `ncu -k regex:...`).  python tools/run_op.py fps 16 16384 1024"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from util import scene_xyz  # noqa: E402
from pdanet_b200 import pointnet2_utils as ops  # noqa: E402

op = sys.argv[1]
a = [float(x) if "." in x else int(x) for x in sys.argv[2:]]
if op == "fps":
    B, N, m = a
    xyz = scene_xyz(N, B, N).cuda()
    for _ in range(3):
        ops.furthest_point_sample(xyz, m)
elif op == "pda_group":
    B, Cc, N, M, r, ns = a
    xyz = scene_xyz(N + M, B, N).cuda()
    feats = torch.randn(B, Cc, N, device="cuda")
    for _ in range(3):
        ops.pda_group(r, ns, xyz, xyz[:, :M].contiguous(), feats)
elif op == "ball_query":
    B, N, M, r, ns = a
    xyz = scene_xyz(N + M, B, N).cuda()
    for _ in range(3):
        ops.ball_query(r, ns, xyz, xyz[:, :M].contiguous())
elif op == "sa_fused":
    import math
    B, N, M, ns, c1, c2, c3 = a
    r = 0.8 if ns > 16 else 0.2
    xyz = scene_xyz(N + M, B, N).cuda()
    feats = torch.rand(B, 1, N, device="cuda")
    dims = [4, c1, c2, c3]
    ws = [torch.randn(dims[i + 1], dims[i], device="cuda") / math.sqrt(dims[i]) for i in range(3)]
    bs = [torch.randn(dims[i + 1], device="cuda") * 0.1 for i in range(3)]
    for _ in range(3):
        ops.sa_fused(r, ns, xyz, xyz[:, :M].contiguous(), feats, ws, bs)
elif op in ("sa_pair", "sa_pair_h"):
    import math
    B, N, M = a
    xyz = scene_xyz(N + M, B, N).cuda()
    feats = torch.rand(B, 1, N, device="cuda")
    ws, bs = [], []
    for dims in ([4, 16, 16, 32], [4, 32, 32, 64]):
        for i in range(3):
            ws.append(torch.randn(dims[i + 1], dims[i], device="cuda") / math.sqrt(dims[i]))
            bs.append(torch.randn(dims[i + 1], device="cuda") * 0.1)
    for _ in range(3):
        ops.sa_fused_pair((0.2, 0.8), (16, 32), xyz, xyz[:, :M].contiguous(), feats, ws, bs, half=op == "sa_pair_h")
elif op == "tc_linear":   # rows k nout npass epilogue  (epilogue: 0 store 1 relu 2 add+LN 3 add+maxpool 4 relu+maxpool)
    from pdanet_b200.tc_linear import PackedLinear
    rows, k, nout, npass, epi = a
    x = torch.randn(rows, k, device="cuda")
    lin = PackedLinear(torch.randn(nout, k, device="cuda") / k ** 0.5, torch.randn(nout, device="cuda"), npass=npass)
    res = torch.randn(rows, nout, device="cuda") if epi in (2, 3) else None
    norm = torch.nn.LayerNorm(nout).cuda() if epi == 2 else None
    for _ in range(3):
        lin(x, epi, residual=res, norm=norm, nsample=32)
elif op == "attention":   # groups ns heads hd
    G, ns, H, hd = a
    qkv = torch.randn(G * ns, 3 * H * hd, device="cuda")
    for _ in range(3):
        ops.group_attention(qkv, ns, H)
torch.cuda.synchronize()
print("done")

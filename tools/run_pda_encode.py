#!/usr/bin/env python
"""One launch config of the fused PDA token encoder for ncu:  python tools/run_pda_encode.py [C] [N] [M] [ns] [radius]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from pdanet_b200 import pointnet2_utils as ops  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 32
r = float(sys.argv[5]) if len(sys.argv) > 5 else 4.8
B = 16
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
xyz = (torch.rand(B, N, 3, generator=g) * torch.tensor([70.0, 80.0, 4.0])).to(dev)
feats_t = torch.randn(B, N, C, generator=g).to(dev)
new_xyz = xyz[:, :M].contiguous()
glob = torch.randn(B * M, C, generator=g).to(dev)
H = C // 2
params = ops.pda_encode_params(torch.randn(H, 12).to(dev) * 0.3, torch.randn(H).to(dev) * 0.1, torch.randn(C, H).to(dev) * 0.2,
                               torch.randn(C).to(dev) * 0.1,
                               [(torch.randn(16, 1).to(dev), torch.rand(16).to(dev)), (torch.randn(8, 16).to(dev), torch.rand(8).to(dev)),
                                (torch.randn(1, 8).abs().to(dev), torch.rand(1).to(dev))],
                               torch.ones(4 * C).to(dev), torch.zeros(4 * C).to(dev))
for _ in range(3):
    y = ops.pda_encode_ln(r, ns, xyz, new_xyz, feats_t, glob, params, 1e-5)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    y = ops.pda_encode_ln(r, ns, xyz, new_xyz, feats_t, glob, params, 1e-5)
e.record()
e.synchronize()
ms = s.elapsed_time(e) / 10
print(f"C {C} N {N} M {M} ns {ns}: {ms:.4f} ms, output {y.numel() * 4 / ms / 1e6:.1f} GB/s")

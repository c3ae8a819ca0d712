#!/usr/bin/env python
"""One launch config of the plain-store GEMM for ncu:  python tools/run_tc_store.py [rows] [k] [nout] [npass] [epi]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from pdanet_b200.tc_linear import PackedLinear  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nout = int(sys.argv[3]) if len(sys.argv) > 3 else 256
npass = int(sys.argv[4]) if len(sys.argv) > 4 else 2
epi = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
x = torch.randn(rows, k, device=dev)
w = torch.randn(nout, k, device=dev) / k ** 0.5
b = torch.randn(nout, device=dev) * 0.1
lin = PackedLinear(w, b, npass=npass)
out = torch.empty(rows, nout, device=dev)
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    lin(x, epi, out=out)
torch.cuda.synchronize()
s.record()
for _ in range(10):
    lin(x, epi, out=out)
t.record()
t.synchronize()
ms = s.elapsed_time(t) / 10
print(f"rows {rows} k {k} nout {nout} npass {npass} epi {epi}: {ms:.4f} ms, {(rows * k + rows * nout) * 4 / ms / 1e6:.0f} GB/s")

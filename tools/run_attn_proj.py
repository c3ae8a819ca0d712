#!/usr/bin/env python
"""One launch config of the fused in_proj + attention kernel (EPI_ATTN) for ncu:  python tools/run_attn_proj.py [rows] [ns] [npass] [f16]
f16 (npass = 4 only): fp16 input and output, i.e. the TMA-fed configuration the model runs."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from pdanet_b200.tc_linear import EPI_ATTN, attn_in_proj  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 32
npass = int(sys.argv[3]) if len(sys.argv) > 3 else 2
f16 = len(sys.argv) > 4 and sys.argv[4] == "f16"
dev = torch.device("cuda:0")
E, heads = 256, 4
x = torch.randn(rows, E, device=dev)
if f16:
    from pdanet_b200.tc_linear import OUT_F16
    x = x.half()
kw = dict(out_fmt=OUT_F16) if f16 else {}
w = torch.randn(3 * E, E, device=dev) / 16
b = torch.randn(3 * E, device=dev) * 0.1
lin = attn_in_proj(w, b, heads, npass=npass)
for _ in range(3):
    y = lin(x, EPI_ATTN, nsample=ns, **kw)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    y = lin(x, EPI_ATTN, nsample=ns, **kw)
e.record()
e.synchronize()
print(f"rows {rows} ns {ns} npass {npass}{' f16' if f16 else ''}: {s.elapsed_time(e) / 10:.4f} ms")
if f16:
    sys.exit(0)

# the same GEMM with a plain STORE epilogue at bn = 192 and bn = 256 (what the attention epilogue costs on top)
from pdanet_b200.tc_linear import EPI_STORE, PackedLinear  # noqa: E402
for bn in (192, 256):
    pl = PackedLinear(w, b, npass=npass, bn=bn)
    out = torch.empty(rows, 3 * E, device=dev)
    for _ in range(3):
        pl(x, EPI_STORE, out=out)
    s.record()
    for _ in range(10):
        pl(x, EPI_STORE, out=out)
    e.record()
    e.synchronize()
    print(f"  STORE bn {bn}: {s.elapsed_time(e) / 10:.4f} ms")

#!/usr/bin/env python
"""Op-level timing of the tcgen05 contractions (csrc/tc_gemm.cu) at the PDA-SSD shapes, beside cuBLAS.

    python tools/bench_tc.py [--iters 20]
Prints one JSON line per shape: ms and TFLOP/s (2*rows*k*nout useful flops) for ours (npass 3 and 1), the
torch 3-pass 3xTF32 composition it replaces (pda_block.Linear3x) and torch's single TF32 / IEEE fp32 matmul.
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from pdanet_b200.tc_linear import PackedLinear, EPI_STORE  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    shapes = [  # (name, rows, k, nout)
        ("L1 in_proj ns32", 524288, 256, 768), ("L1 out_proj ns32", 524288, 256, 256),
        ("L1 lin1 ns32", 524288, 256, 128), ("L1 lin2 ns32", 524288, 128, 256),
        ("L2 in_proj ns32", 262144, 512, 1536), ("L2 out_proj ns32", 262144, 512, 512),
        ("L2 lin1 ns32", 262144, 512, 256), ("L2 lin2 ns32", 262144, 256, 512),
        ("L5 mlp2 ns32", 131072, 256, 512), ("L5 mlp3 ns32", 131072, 512, 1024),
    ]
    for name, rows, k, nout in shapes:
        x = torch.randn(rows, k, device=dev)
        w = torch.randn(nout, k, device=dev) / k ** 0.5
        b = torch.randn(nout, device=dev)
        out = torch.empty(rows, nout, device=dev)
        flops = 2.0 * rows * k * nout
        res = {"shape": name, "rows": rows, "k": k, "nout": nout}
        for npass in (3, 2, 1):
            lin = PackedLinear(w, b, npass=npass)
            ms = timeit(lambda: lin(x, EPI_STORE, out=out), args.iters)
            res[f"tc{npass}_ms"] = round(ms, 4)
            res[f"tc{npass}_tflops"] = round(flops / ms / 1e9, 1)
        from pdanet_b200 import _lib
        _lib.lib().pdab_set_cta_pairs(0)
        lin = PackedLinear(w, b, npass=3)
        res["tc3_single_cta_ms"] = round(timeit(lambda: lin(x, EPI_STORE, out=out), args.iters), 4)
        _lib.lib().pdab_set_cta_pairs(1)
        wt = w.t().contiguous()
        hi = (wt.view(torch.int32) & -8192).view(torch.float32)
        lo = wt - hi

        def torch3():
            xh = (x.view(torch.int32) & -8192).view(torch.float32)
            xl = x - xh
            y = torch.addmm(b, xh, hi)
            y.addmm_(xh, lo)
            y.addmm_(xl, hi)
            return y
        torch.backends.cuda.matmul.allow_tf32 = True
        res["torch3x_ms"] = round(timeit(torch3, args.iters), 4)
        res["torch_tf32_ms"] = round(timeit(lambda: torch.addmm(b, x, wt), args.iters), 4)
        torch.backends.cuda.matmul.allow_tf32 = False
        res["torch_fp32_ms"] = round(timeit(lambda: torch.addmm(b, x, wt), args.iters), 4)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

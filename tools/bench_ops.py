#!/usr/bin/env python
"""Op-level sweep (BASELINE.json configs[3]): our kernels vs the reference's CUDA kernels rebuilt for
sm_100a (oracle/_ref), same inputs, CUDA-event timing, outputs compared bit-for-bit where both run.

    python tools/bench_ops.py [--quick] > gpurun_out/ops_sweep.jsonl
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_ref_iou3d, load_ref_pointnet2  # noqa: E402
from util import random_boxes, scene_xyz  # noqa: E402
from pdanet_b200 import iou3d_nms_utils, pointnet2_utils as ops  # noqa: E402


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)


def vp(t):
    return C.c_void_p(t.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--fps-only", action="store_true")
    args = ap.parse_args()
    pn, iou = load_ref_pointnet2(), load_ref_iou3d()
    B = 16

    fps_cases = [(16384, 4096), (4096, 1024), (16384, 1024), (16384, 512)]
    if not args.quick:
        fps_cases += [(65536, 4096), (65536, 1024), (262144, 512)]
    for N, m in fps_cases:
        b = B if N <= 16384 else 4
        xyz = scene_xyz(N, b, N).cuda()
        ours = lambda: ops.furthest_point_sample(xyz, m)
        t_min, t_avg = timeit(ours)
        rec = {"op": "fps", "B": b, "N": N, "m": m, "ours_ms": round(t_min, 4), "ours_avg_ms": round(t_avg, 4),
               "updates_per_s": b * N * (m - 1) / (t_min / 1e3)}
        if pn is not None:
            temp = torch.empty(b, N, device="cuda")
            idx = torch.zeros(b, m, dtype=torch.int32, device="cuda")

            def ref():
                temp.fill_(1e10)
                pn.ref_fps(b, N, m, vp(xyz), vp(temp), vp(idx))
            r_min, _ = timeit(ref, iters=3, warmup=1)
            rec.update(ref_ms=round(r_min, 4), speedup=round(r_min / t_min, 2), equal=bool(torch.equal(idx, ours())))
        print(json.dumps(rec), flush=True)

    if args.fps_only:
        return
    for N, M, r, ns in [(16384, 4096, 0.2, 16), (16384, 4096, 0.8, 32), (4096, 1024, 1.6, 32), (1024, 512, 4.8, 32),
                        (512, 256, 6.4, 32)] + ([] if args.quick else [(65536, 16384, 0.8, 32), (16384, 4096, 4.8, 64)]):
        xyz = scene_xyz(N + M, B, N).cuda()
        new_xyz = xyz[:, :M].contiguous()
        ours = lambda: ops.ball_query(r, ns, xyz, new_xyz)
        t_min, _ = timeit(ours)
        rec = {"op": "ball_query", "B": B, "N": N, "M": M, "radius": r, "nsample": ns, "ours_ms": round(t_min, 4),
               "pairs_per_s": B * N * M / (t_min / 1e3)}
        if pn is not None:
            idx = torch.zeros(B, M, ns, dtype=torch.int32, device="cuda")

            def ref():
                idx.zero_()
                pn.ref_ball_query(B, N, M, C.c_float(r), ns, vp(new_xyz), vp(xyz), vp(idx))
            r_min, _ = timeit(ref, iters=3, warmup=1)
            rec.update(ref_ms=round(r_min, 4), speedup=round(r_min / t_min, 2), equal=bool(torch.equal(idx, ours())))
        print(json.dumps(rec), flush=True)

    for Cc, N, M, ns in [(3, 16384, 4096, 32), (64, 4096, 1024, 32), (128, 1024, 512, 32), (256, 512, 256, 32)]:
        feats = torch.randn(B, Cc, N, device="cuda")
        idx = torch.randint(0, N, (B, M, ns), dtype=torch.int32, device="cuda")
        ours = lambda: ops.grouping_operation(feats, idx)
        t_min, _ = timeit(ours)
        nbytes = B * (2 * 4 * Cc * M * ns + 4 * M * ns)
        rec = {"op": "group_points", "B": B, "C": Cc, "N": N, "M": M, "nsample": ns, "ours_ms": round(t_min, 4),
               "GBps": round(nbytes / t_min / 1e6, 1)}
        if pn is not None:
            out = torch.empty(B, Cc, M, ns, device="cuda")
            ref = lambda: pn.ref_group(B, Cc, N, M, ns, vp(feats), vp(idx), vp(out))
            r_min, _ = timeit(ref, iters=3, warmup=1)
            rec.update(ref_ms=round(r_min, 4), speedup=round(r_min / t_min, 2), equal=bool(torch.equal(out, ours())))
        print(json.dumps(rec), flush=True)

    for Cc, N, M, r, ns in [(64, 4096, 1024, 0.8, 16), (64, 4096, 1024, 1.6, 32), (128, 1024, 512, 1.6, 16),
                            (128, 1024, 512, 4.8, 32)]:
        xyz = scene_xyz(N + M, B, N).cuda()
        new_xyz = xyz[:, :M].contiguous()
        feats = torch.randn(B, Cc, N, device="cuda")
        t_min, _ = timeit(lambda: ops.pda_group(r, ns, xyz, new_xyz, feats))
        nbytes = B * (12 * N + 4 * Cc * N + 12 * M + 4 * (7 + Cc) * M * ns)
        print(json.dumps({"op": "pda_group", "B": B, "C": Cc, "N": N, "M": M, "radius": r, "nsample": ns,
                          "ours_ms": round(t_min, 4), "GBps": round(nbytes / t_min / 1e6, 1)}), flush=True)

    for n, thresh in [(256, 0.01), (1024, 0.1), (4096, 0.1)]:
        boxes = random_boxes(n, n, extent=(70.0, 80.0, 2.0)).cuda()
        scores = torch.rand(n, device="cuda")
        t_min, _ = timeit(lambda: iou3d_nms_utils.nms_gpu(boxes, scores, thresh))
        rec = {"op": "nms_gpu (host keep list, 1 scene)", "n": n, "thresh": thresh, "ours_ms": round(t_min, 4)}
        if iou is not None:
            order = scores.sort(0, descending=True)[1]
            sb = boxes[order].contiguous()
            keep = torch.zeros(n, dtype=torch.int64)
            r_min, _ = timeit(lambda: iou.nms_gpu(sb, keep, thresh), iters=3, warmup=1)
            rec.update(ref_ms=round(r_min, 4), speedup=round(r_min / t_min, 2))
        print(json.dumps(rec), flush=True)
        sb16 = boxes[scores.sort(0, descending=True)[1]].unsqueeze(0).repeat(16, 1, 1).contiguous()
        counts = torch.full((16,), n, dtype=torch.int32, device="cuda")
        t_min, _ = timeit(lambda: iou3d_nms_utils.nms_batched(sb16, counts, thresh))
        print(json.dumps({"op": "nms_batched (device, 16 scenes)", "n": n, "ours_ms": round(t_min, 4),
                          "pairs_per_s": 16 * n * (n - 1) / 2 / (t_min / 1e3)}), flush=True)


if __name__ == "__main__":
    main()

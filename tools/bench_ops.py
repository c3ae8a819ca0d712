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

    # F-FPS on a precomputed distance matrix (PB/src/sampling_gpu.cu:256-416); the matrix is B x N x N fp32, which bounds N
    for N, m in [(2048, 512), (4096, 1024)] + ([] if args.quick else [(8192, 1024)]):
        b = 16 if N <= 4096 else 4
        g = torch.Generator().manual_seed(N)
        pts = torch.randn(b, N, 8, generator=g).cuda()
        dist = torch.cdist(pts, pts).pow(2).contiguous()
        ours = lambda: ops.furthest_point_sample_with_dist(dist, m)
        t_min, t_avg = timeit(ours)
        rec = {"op": "f_fps (distance matrix)", "B": b, "N": N, "m": m, "ours_ms": round(t_min, 4),
               "updates_per_s": b * N * (m - 1) / (t_min / 1e3), "matrix_GB": round(b * N * N * 4 / 1e9, 2)}
        if pn is not None:
            temp = torch.empty(b, N, device="cuda")
            idx = torch.zeros(b, m, dtype=torch.int32, device="cuda")

            def ref():
                temp.fill_(1e10)
                pn.ref_fps_with_dist(b, N, m, vp(dist), vp(temp), vp(idx))
            r_min, _ = timeit(ref, iters=3, warmup=1)
            rec.update(ref_ms=round(r_min, 4), speedup=round(r_min / t_min, 2), equal=bool(torch.equal(idx, ours())))
        print(json.dumps(rec), flush=True)
        del dist

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
        # the pybind-level call both sides: sorted device boxes in, host keep list out (IOU/src/iou3d_nms.cpp:90-136)
        from pdanet_b200 import iou3d_nms_cuda as shim
        order = scores.sort(0, descending=True)[1]
        sb = boxes[order].contiguous()
        keep_o = torch.zeros(n, dtype=torch.int64)
        t_min, _ = timeit(lambda: shim.nms_gpu(sb, keep_o, thresh), iters=10, warmup=3)
        w_min, _ = timeit(lambda: iou3d_nms_utils.nms_gpu(boxes, scores, thresh), iters=10, warmup=3)
        rec = {"op": "nms_gpu (pybind call: host keep list, 1 scene)", "n": n, "thresh": thresh, "ours_ms": round(t_min, 4),
               "ours_python_wrapper_ms": round(w_min, 4)}
        if iou is not None:
            keep = torch.zeros(n, dtype=torch.int64)
            r_min, _ = timeit(lambda: iou.nms_gpu(sb, keep, thresh), iters=10, warmup=3)
            num_o = shim.nms_gpu(sb, keep_o, thresh)
            num_r = iou.nms_gpu(sb, keep, thresh)
            rec.update(ref_ms=round(r_min, 4), speedup=round(r_min / t_min, 2),
                       equal=bool(num_o == num_r and torch.equal(keep_o[:num_o], keep[:num_r])))
        print(json.dumps(rec), flush=True)
        sb16 = boxes[scores.sort(0, descending=True)[1]].unsqueeze(0).repeat(16, 1, 1).contiguous()
        counts = torch.full((16,), n, dtype=torch.int32, device="cuda")
        t_min, _ = timeit(lambda: iou3d_nms_utils.nms_batched(sb16, counts, thresh))
        print(json.dumps({"op": "nms_batched (device, 16 scenes)", "n": n, "ours_ms": round(t_min, 4),
                          "pairs_per_s": 16 * n * (n - 1) / 2 / (t_min / 1e3)}), flush=True)


def fused_sa_rows(pn):
    """BASELINE configs[3] "fused group+MLP+max-pool per SA layer vs reference ops": our fused layer (one or a few launches, no
    grouped tensor) against the reference's unfused chain on the same module parameters — the reference CUDA kernels rebuilt
    for sm_100a (ball query + 2 x group) + torch sub / cat / Conv2d+BN+ReLU x 3 / max_pool2d, PB/pointnet2_utils.py:689-704,
    PB/pointnet2_modules.py:1655-1674."""
    import torch.nn.functional as F
    from pdanet_b200.pointnet2_modules import PointnetSAModuleMSG_WithSampling
    if pn is None:
        return
    B = 16
    shapes = [("kitti_L0", 16384, 4096, 1, [0.2, 0.8], [16, 32], [[1, 16, 16, 32], [1, 32, 32, 64]], [64]),
              ("kitti_L5", 512, 256, 256, [4.8, 6.4], [16, 32], [[256, 256, 256, 512], [256, 256, 512, 1024]], [512]),
              ("once_L5", 2048, 1024, 256, [4.8, 8.4, 12.8], [16, 32, 64],
               [[256, 256, 256, 512], [256, 256, 256, 512], [256, 256, 512, 512]], [512])]
    for tag, N, M, Cin, radii, nsamples, mlps, agg in shapes:
        b = B if tag != "once_L5" else 8
        torch.manual_seed(0)
        mod = PointnetSAModuleMSG_WithSampling(npoint_list=[M], sample_range_list=[-1], sample_type_list=["D-FPS"], radii=radii,
                                               nsamples=nsamples, mlps=mlps, aggregation_mlp=agg, confidence_mlp=[],
                                               num_class=3).cuda().eval()
        xyz = scene_xyz(N + M, b, N).cuda()
        new_xyz = xyz[:, :M].contiguous()
        feats = torch.rand(b, Cin, N, device="cuda")
        with torch.no_grad():
            ours = lambda: mod(xyz, feats, None, ctr_xyz=new_xyz)[1]
            t_min, _ = timeit(ours, iters=10, warmup=3)

            def ref():
                xyz_t = xyz.transpose(1, 2).contiguous()
                outs = []
                for i, (r, ns) in enumerate(zip(radii, nsamples)):
                    idx = torch.zeros(b, M, ns, dtype=torch.int32, device="cuda")
                    pn.ref_ball_query(b, N, M, C.c_float(r), ns, vp(new_xyz), vp(xyz), vp(idx))
                    gx = torch.empty(b, 3, M, ns, device="cuda")
                    pn.ref_group(b, 3, N, M, ns, vp(xyz_t), vp(idx), vp(gx))
                    gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
                    gf = torch.empty(b, Cin, M, ns, device="cuda")
                    pn.ref_group(b, Cin, N, M, ns, vp(feats), vp(idx), vp(gf))
                    y = mod.mlps[i](torch.cat([gx, gf], dim=1))
                    outs.append(F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1))
                return mod.aggregation_layer(torch.cat(outs, dim=1))
            r_min, _ = timeit(ref, iters=3, warmup=1)
            a, c = ours(), ref()
            err = float((a - c).abs().max() / c.abs().max())
        flops = 2.0 * b * M * sum(ns * sum(m[k] * m[k + 1] + (3 * m[1] if k == 0 else 0) for k in range(len(m) - 1))
                                  for ns, m in zip(nsamples, mlps))
        print(json.dumps({"op": "fused SA layer (ball query + group + MLP + max-pool + aggregation) vs reference chain",
                          "layer": tag, "B": b, "N": N, "M": M, "C": Cin, "radii": radii, "nsample": nsamples,
                          "ours_ms": round(t_min, 4), "ref_ms": round(r_min, 4), "speedup": round(r_min / t_min, 2),
                          "ours_TFLOPs": round(flops / t_min / 1e9, 1), "max_rel_diff": err,
                          "tc_passes": mod.tc_passes}), flush=True)


if __name__ == "__main__":
    main()
    fused_sa_rows(load_ref_pointnet2())

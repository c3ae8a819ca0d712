"""Generates the golden vectors in tests/golden/ by running the REFERENCE's own CUDA kernels
(oracle/_ref/*.so, built unmodified from /root/reference by oracle/build_ref.py) on a B200.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/            # then commit

The reference ships no golden vectors, known-answer tests or fixtures of its own;
these files are outputs of the reference itself and pin both the CPU oracle (tests/test_oracle_cpu.py)
and the CUDA kernels (tests/test_gpu_ops.py).  Inputs are seeded (tests/util.py)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_ref_iou3d, load_ref_pointnet2  # noqa: E402
from util import adversarial_boxes, random_boxes, scene_xyz  # noqa: E402

OUT = ROOT / "gpurun_out" / "golden"


def vp(t):
    return C.c_void_p(t.data_ptr())


def main():
    assert torch.cuda.is_available()
    pn = load_ref_pointnet2()
    iou = load_ref_iou3d()
    assert pn is not None and iou is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    OUT.mkdir(parents=True, exist_ok=True)

    # ---- FPS (farthest_point_sampling_kernel, PB/src/sampling_gpu.cu:93-253)
    fps = {}
    for tag, B, N, m, kw in [("rand2048", 2, 2048, 256, {}), ("dups2048", 2, 2048, 256, dict(duplicate_frac=0.3)),
                             ("grid1500", 1, 1500, 200, dict(quantize=2.0)), ("grid1000", 1, 1000, 128, dict(quantize=2.0)),
                             ("tiny37", 2, 37, 37, dict(quantize=4.0)), ("rand4096", 1, 4096, 512, {})]:
        xyz = scene_xyz(100 + N, B, N, **kw)
        x = xyz.cuda()
        temp = torch.full((B, N), 1e10, device="cuda")
        idx = torch.zeros(B, m, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        assert pn.ref_fps(B, N, m, vp(x), vp(temp), vp(idx)) == 0
        fps[f"xyz_{tag}"] = xyz.numpy()
        fps[f"idx_{tag}"] = idx.cpu().numpy()
    np.savez_compressed(OUT / "ref_fps.npz", **fps)

    # ---- ball query + group (PB/src/ball_query_gpu.cu:9-67, PB/src/group_points_gpu.cu:53-92)
    bg = {}
    for tag, B, N, M, r, ns, kw in [("r0.8", 2, 2048, 256, 0.8, 16, dict(hi=(20.0, 20.0, 1.0), lo=(0.0, 0.0, -1.0))),
                                    ("r2_grid", 1, 1500, 130, 2.0, 8, dict(quantize=1.0)),
                                    ("r4.8", 2, 512, 128, 4.8, 32, {})]:
        xyz = scene_xyz(200 + N, B, N, **kw)
        new = xyz[:, :M].contiguous()
        new[:, -1] = 1e4
        feat = torch.randn(B, 4, N, generator=torch.Generator().manual_seed(N))
        x, c, f = xyz.cuda(), new.cuda(), feat.cuda()
        idx = torch.zeros(B, M, ns, dtype=torch.int32, device="cuda")
        out = torch.zeros(B, 4, M, ns, device="cuda")
        torch.cuda.synchronize()
        assert pn.ref_ball_query(B, N, M, C.c_float(r), ns, vp(c), vp(x), vp(idx)) == 0
        assert pn.ref_group(B, 4, N, M, ns, vp(f), vp(idx), vp(out)) == 0
        bg.update({f"xyz_{tag}": xyz.numpy(), f"new_{tag}": new.numpy(), f"feat_{tag}": feat.numpy(),
                   f"radius_{tag}": np.float32(r), f"nsample_{tag}": np.int32(ns),
                   f"idx_{tag}": idx.cpu().numpy(), f"grouped_{tag}": out.cpu().numpy()})
    np.savez_compressed(OUT / "ref_ball_group.npz", **bg)

    # ---- rotated IoU + NMS (IOU/src/iou3d_nms_kernel.cu:236-311, IOU/src/iou3d_nms.cpp:90-136)
    nm = {}
    for tag, boxes, thresh in [("kitti256", random_boxes(1, 256), 0.01), ("dense300", random_boxes(2, 300, extent=(20.0, 20.0, 1.0)), 0.1),
                               ("adv256", adversarial_boxes(3, 256), 0.1)]:
        n = boxes.shape[0]
        b = boxes.cuda()
        m = torch.zeros(n, n, device="cuda")
        iou.boxes_iou_bev_gpu(b, b, m)
        keep = torch.zeros(n, dtype=torch.int64)
        num = iou.nms_gpu(b, keep, thresh)
        m = m.cpu()
        nm.update({f"boxes_{tag}": boxes.numpy(), f"thresh_{tag}": np.float32(thresh), f"iou_{tag}": m.numpy(),
                   f"keep_{tag}": keep[:num].numpy(),
                   f"clear_{tag}": np.bool_(not bool(((torch.nan_to_num(m) - thresh).abs() < 1e-4).any()))})
    np.savez_compressed(OUT / "ref_nms.npz", **nm)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()

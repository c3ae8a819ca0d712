"""CPU tests of the oracle itself: pinned against the reference where the reference runs here,
against committed golden vectors produced by the reference's CUDA kernels on a B200
(tests/golden/make_golden.py), and against independent restatements."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ops
from conftest import GOLDEN
from util import adversarial_boxes, random_boxes, scene_xyz


# ------------------------------------------------------------------ IoU / NMS pinned on the reference CPU code

@pytest.mark.parametrize("maker,seed,n", [(random_boxes, 0, 300), (random_boxes, 1, 257), (adversarial_boxes, 2, 256)])
def test_iou_bev_bit_exact_vs_reference_cpu(ref_iou3d, maker, seed, n):
    """oracle/pdab_oracle.c::orc_iou_bev == reference boxes_iou_bev_cpu (IOU/src/iou3d_cpu.cpp:232-252), bit for bit."""
    boxes = maker(seed, n)
    want = torch.zeros(n, n)
    got = torch.zeros(n, n)
    ref_iou3d.boxes_iou_bev_cpu(boxes, boxes, want)
    oracle.boxes_iou_bev_cpu(boxes, boxes, got)
    assert torch.equal(torch.nan_to_num(want, nan=-1.0), torch.nan_to_num(got, nan=-1.0))
    assert (want > 0).sum() > n  # the case is not vacuous


def test_nms_matches_greedy_on_reference_cpu_iou(ref_iou3d):
    """The oracle's tiled-bitmask NMS equals plain greedy suppression driven by the reference's CPU IoU."""
    n, thresh = 500, 0.1
    boxes = random_boxes(5, n, extent=(25.0, 25.0, 1.0))
    scores = torch.rand(n, generator=torch.Generator().manual_seed(6))
    order = scores.sort(dim=0, descending=True, stable=True)[1]
    sb = boxes[order].contiguous()
    iou = torch.zeros(n, n)
    ref_iou3d.boxes_iou_bev_cpu(sb, sb, iou)
    removed = np.zeros(n, bool)
    want = []
    for i in range(n):
        if not removed[i]:
            want.append(i)
            removed[i + 1:] |= (iou[i, i + 1:] > thresh).numpy()
    keep = torch.zeros(n, dtype=torch.int64)
    num = oracle.nms_gpu(sb, keep, thresh)
    assert keep[:num].tolist() == want
    assert 0 < num < n


def test_nms_edge_cases():
    keep = torch.zeros(1, dtype=torch.int64)
    assert oracle.nms_gpu(torch.zeros(0, 7), torch.zeros(0, dtype=torch.int64), 0.1) == 0
    assert oracle.nms_gpu(random_boxes(0, 1), keep, 0.1) == 1 and keep[0] == 0
    same = random_boxes(3, 1).repeat(130, 1).contiguous()  # coincident boxes across tile borders
    keep = torch.zeros(130, dtype=torch.int64)
    assert oracle.nms_gpu(same, keep, 0.5) == 1


# ------------------------------------------------------------------ FPS: literal tree vs closed-form tie rule

def _bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def fps_closed_form(xyz: np.ndarray, m: int) -> np.ndarray:
    """FPS with the tie rule SURVEY.md §8a-1 derives: among maxima pick argmin (bitrev_L(k mod BS), k).
    Only exact when every distance is exactly representable (grid-snapped inputs)."""
    n = xyz.shape[0]
    bs = oracle.lib().orc_opt_n_threads(n)
    L = int(math.log2(bs))
    key = np.array([(_bitrev(k % bs, L) << 32) | k for k in range(n)], dtype=np.int64)
    temp = np.full(n, 1e10, np.float32)
    out = [0]
    for _ in range(1, m):
        d = ((xyz - xyz[out[-1]]) ** 2).sum(1).astype(np.float32)
        temp = np.minimum(temp, d)
        ties = np.flatnonzero(temp == temp.max())
        out.append(int(ties[np.argmin(key[ties])]))
    return np.array(out, np.int32)


@pytest.mark.parametrize("n,m", [(2048, 256), (1000, 100), (1500, 64), (300, 300), (64, 16), (5, 5)])
def test_fps_tree_emulation_equals_bitrev_rule_on_tie_heavy_clouds(n, m):
    xyz = scene_xyz(11 + n, 1, n, quantize=2.0, duplicate_frac=0.1)  # grid-snapped: exact distances, many ties
    got = oracle.fps(xyz, m)[0].numpy()
    want = fps_closed_form(xyz[0].numpy(), m)
    assert np.array_equal(got, want)


def test_fps_basic_properties():
    xyz = scene_xyz(3, 2, 4096)
    idx = oracle.fps(xyz, 512)
    assert idx.dtype == torch.int32 and idx.shape == (2, 512)
    assert (idx[:, 0] == 0).all()
    for b in range(2):
        assert len(set(idx[b].tolist())) == 512  # distinct points never repeat
    # the greedy min-distance sequence is non-increasing
    p = xyz[0][idx[0].long()]
    d = torch.cdist(p.double(), p.double())
    mins = [d[j, :j].min().item() for j in range(1, 512)]
    assert all(mins[j] >= mins[j + 1] - 1e-6 for j in range(len(mins) - 1))


def test_fps_with_dist_equals_fps_on_the_same_distances():
    xyz = scene_xyz(8, 1, 512, quantize=1.0)
    d = ((xyz[:, :, None, :] - xyz[:, None, :, :]) ** 2).sum(-1).contiguous()  # exact on the grid
    assert torch.equal(oracle.fps_with_dist(d, 64), oracle.fps(xyz, 64))


# ------------------------------------------------------------------ ball query / group / gather

def test_ball_query_semantics():
    xyz = scene_xyz(4, 2, 1024, quantize=0.5)
    new_xyz = xyz[:, :64].contiguous()
    r, ns = 3.0, 8
    idx = oracle.ball_query(r, ns, xyz, new_xyz)
    d2 = ((new_xyz[:, :, None, :] - xyz[:, None, :, :]) ** 2).sum(-1)  # exact on the grid
    for b in range(2):
        for j in range(64):
            hits = torch.nonzero(d2[b, j] < np.float32(r) * np.float32(r)).flatten().tolist()
            want = hits[:ns] + [hits[0]] * (ns - len(hits[:ns])) if hits else [0] * ns
            assert idx[b, j].tolist() == want
    far = torch.full((1, 3, 3), 1e4)
    assert (oracle.ball_query(0.5, 4, xyz[:1], far) == 0).all()  # empty balls keep the zero row


def test_ball_query_dilated_emits_zero_distance_point_twice():
    xyz = torch.tensor([[[0., 0, 0], [1, 0, 0], [3, 0, 0]]])
    idx = oracle.ball_query_dilated(2.0, 0.0, 4, xyz, xyz[:, :1].contiguous())
    assert idx[0, 0].tolist() == [0, 0, 1, 0]  # k=0 twice (d2==0 and in [0,4)), then k=1, then first-hit fill


def test_group_and_gather_are_index_selects():
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(2, 5, 100, generator=g)
    idx2 = torch.randint(0, 100, (2, 7), generator=g, dtype=torch.int32)
    idx3 = torch.randint(0, 100, (2, 7, 3), generator=g, dtype=torch.int32)
    assert torch.equal(oracle.gather(feats, idx2), torch.gather(feats, 2, idx2.long()[:, None].expand(-1, 5, -1)))
    want = torch.gather(feats, 2, idx3.long().view(2, 1, 21).expand(-1, 5, -1)).view(2, 5, 7, 3)
    assert torch.equal(oracle.group(feats, idx3), want)
    grad = torch.zeros(2, 5, 100)
    oracle.group_points_grad_wrapper(2, 5, 100, 7, 3, want.contiguous(), idx3, grad)
    ref = torch.zeros(2, 5, 100).scatter_add_(2, idx3.long().view(2, 1, 21).expand(-1, 5, -1), want.view(2, 5, 21))
    assert torch.allclose(grad, ref, atol=1e-5)


# ------------------------------------------------------------------ fused-op restatements vs torch

def test_topk_ctr_is_a_valid_topk_of_the_sigmoid_scores():
    g = torch.Generator().manual_seed(1)
    cls = torch.randn(3, 777, 3, generator=g) * 4
    cls[:, ::5] = cls[:, 1::5][:, : cls[:, ::5].shape[1]]  # exact ties
    idx = oracle.topk_ctr(cls, 300).long()
    score = torch.sigmoid(cls.max(-1)[0])
    want = torch.topk(score, 300, dim=-1)[0]
    assert torch.equal(torch.gather(score, 1, idx), want)
    mx = cls.max(-1)[0]
    picked = torch.gather(mx, 1, idx)
    for b in range(3):
        assert len(set(idx[b].tolist())) == 300
        v, i = picked[b].tolist(), idx[b].tolist()
        assert all(v[q] > v[q + 1] or (v[q] == v[q + 1] and i[q] < i[q + 1]) for q in range(299))


def test_pda_group_matches_torch_restatement():
    xyz = scene_xyz(9, 2, 512, hi=(10.0, 10.0, 1.0), lo=(0.0, 0.0, -1.0))
    feats = torch.randn(2, 6, 512, generator=torch.Generator().manual_seed(2))
    new_xyz = xyz[:, :40].contiguous()
    out, idx = oracle.pda_group(1.6, 16, xyz, new_xyz, feats)
    want = torch_ops.QueryAndGroup_alone_grouped_density_directional(1.6, 16)(xyz, new_xyz, feats)
    assert out.shape == (2, 13, 40, 16)
    assert torch.equal(out[:, :3], want[:, :3]) and torch.equal(out[:, 7:], want[:, 7:])
    assert torch.allclose(out[:, 3:7], want[:, 3:7], rtol=1e-5, atol=1e-7)


def test_sa_mlp_maxpool_matches_torch_unfused_path():
    g = torch.Generator().manual_seed(3)
    xyz = scene_xyz(10, 2, 600, hi=(8.0, 8.0, 1.0), lo=(0.0, 0.0, -1.0))
    feats = torch.randn(2, 1, 600, generator=g)
    new_xyz = xyz[:, :50].contiguous()
    dims = [4, 16, 16, 32]
    ws = [torch.randn(dims[i + 1], dims[i], generator=g) * 0.5 for i in range(3)]
    bs = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(3)]
    got = oracle.sa_mlp_maxpool(0.8, 16, xyz, new_xyz, feats, ws, bs)
    x = torch_ops.QueryAndGroup(0.8, 16)(xyz, new_xyz, feats)
    for w, b in zip(ws, bs):
        x = torch.relu(torch.einsum("oc,bcms->boms", w, x) + b[None, :, None, None])
    assert torch.allclose(got, x.max(-1)[0], rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ golden vectors from the reference CUDA kernels

def _golden(name):
    path = GOLDEN / name
    if not path.exists():
        pytest.skip(f"{name} not generated yet (tests/golden/make_golden.py runs on the GPU box)")
    return np.load(path)


def test_oracle_equals_reference_cuda_fps_golden():
    z = _golden("ref_fps.npz")
    for tag in [k[4:] for k in z.files if k.startswith("xyz_")]:
        xyz = torch.from_numpy(z[f"xyz_{tag}"])
        want = z[f"idx_{tag}"]
        got = oracle.fps(xyz, want.shape[1]).numpy()
        assert np.array_equal(got, want), tag


def test_oracle_equals_reference_cuda_ball_query_group_golden():
    z = _golden("ref_ball_group.npz")
    for tag in [k[4:] for k in z.files if k.startswith("xyz_")]:
        xyz, new_xyz = torch.from_numpy(z[f"xyz_{tag}"]), torch.from_numpy(z[f"new_{tag}"])
        r, ns = float(z[f"radius_{tag}"]), int(z[f"nsample_{tag}"])
        idx = oracle.ball_query(r, ns, xyz, new_xyz)
        assert np.array_equal(idx.numpy(), z[f"idx_{tag}"]), tag
        feats = torch.from_numpy(z[f"feat_{tag}"])
        assert np.array_equal(oracle.group(feats, idx).numpy(), z[f"grouped_{tag}"]), tag


def test_oracle_nms_vs_reference_cuda_nms_golden():
    """The CPU oracle has no FMA contraction and uses libm, the GPU reference does / uses libdevice: IoU values agree
    to ~1e-6 and keep lists are identical on inputs whose IoUs stay clear of the threshold (recorded margin)."""
    z = _golden("ref_nms.npz")
    for tag in [k[6:] for k in z.files if k.startswith("boxes_")]:
        boxes = torch.from_numpy(z[f"boxes_{tag}"])
        thresh = float(z[f"thresh_{tag}"])
        iou = torch.zeros(boxes.shape[0], boxes.shape[0])
        oracle.boxes_iou_bev_cpu(boxes, boxes, iou)
        ref_iou = torch.from_numpy(z[f"iou_{tag}"])
        assert torch.allclose(iou, ref_iou, atol=2e-5), tag
        if bool(z[f"clear_{tag}"]):
            keep = torch.zeros(boxes.shape[0], dtype=torch.int64)
            num = oracle.nms_gpu(boxes, keep, thresh)
            assert keep[:num].tolist() == z[f"keep_{tag}"].tolist(), tag


def test_three_nn_interpolate_and_points_in_boxes_oracle():
    """Oracle restatements of the §8f-3/4 ops against straightforward float64 numpy/torch statements."""
    g = torch.Generator().manual_seed(0)
    unknown, known = torch.rand(2, 200, 3, generator=g) * 10, torch.rand(2, 50, 3, generator=g) * 10
    d2, idx = oracle.three_nn(unknown, known)
    full = torch.cdist(unknown.double(), known.double()) ** 2
    want_d, want_i = full.topk(3, dim=2, largest=False)
    assert torch.equal(idx.long(), want_i)                       # random clouds: no ties
    assert torch.allclose(d2.double(), want_d, rtol=1e-5, atol=1e-6)
    feats = torch.randn(2, 7, 50, generator=g)
    w = torch.rand(2, 200, 3, generator=g)
    out = oracle.three_interpolate(feats, idx, w)
    gathered = torch.stack([torch.gather(feats, 2, idx[:, :, k].long().unsqueeze(1).expand(2, 7, 200)) for k in range(3)], -1)
    assert torch.allclose(out.double(), (gathered.double() * w.unsqueeze(1).double()).sum(-1), rtol=1e-5, atol=1e-6)
    # fewer than three candidates -> (+inf, 0) padding, as (float)1e40 in the reference
    d2, idx = oracle.three_nn(unknown[:, :5], known[:, :2])
    assert torch.isinf(d2[..., 2]).all() and (idx[..., 2] == 0).all()
    # points in boxes: axis-aligned and rotated box, first match wins
    boxes = torch.tensor([[[0.0, 0.0, 0.0, 4.0, 2.0, 2.0, 0.0], [0.0, 0.0, 0.0, 4.0, 2.0, 2.0, 1.5707964]]])
    pts = torch.tensor([[[1.9, 0.9, 0.9], [0.5, 1.9, 0.0], [3.0, 0.0, 0.0], [0.0, 0.0, 1.1]]])
    assert oracle.points_in_boxes(pts, boxes)[0].tolist() == [0, 1, -1, -1]


# ------------------------------------------------------------------ the acceptance rule of the batched FPS kernel, on the host

@pytest.mark.parametrize("n,m,P,NB,T,cloud", [
    (4096, 1024, 8, 1, 512, 0),      # KITTI L1 shape, LiDAR-like slab
    (16384, 700, 16, 2, 512, 1),     # KITTI L0 layout, dense clusters + background
    (2048, 300, 4, 1, 512, 2),       # all points equal: every round falls back to the plain argmax
    (4096, 500, 8, 1, 512, 3),       # lattice: masses of exactly equal distances (overflowing lists)
    (1000, 999, 2, 1, 512, 0),       # ragged size, nearly every point sampled
])
def test_fps_rounds_acceptance_rule_replayed_on_host(tmp_path, n, m, P, NB, T, cloud):
    """csrc/fps_pruned.cu takes several samples per barrier: candidates above a threshold are ranked and the longest prefix in
    which no candidate is moved by an earlier one, nor overtaken by a point hidden behind a thread's best, is accepted.
    tools/fps_batch_sim.c replays exactly that rule (same keys, same fp32 distance expression, same bucket -> thread layout,
    same threshold controller) on the host and compares idx AND temp with the oracle's literal emulation of the reference
    kernel; it exits non-zero on any mismatch."""
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "fps_batch_sim"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tools" / "fps_batch_sim.c"),
                           str(ROOT / "oracle" / "pdab_oracle.c"), "-lm"])
    out = subprocess.run([str(exe), str(n), str(m), str(P), str(NB), str(T), str(cloud), "7", "10", "28"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "idx mismatches 0, temp mismatches 0" in out.stdout

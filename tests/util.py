"""Seeded input generators shared by the CPU and GPU tests."""
import math

import torch


def scene_xyz(seed, b, n, lo=(0.0, -40.0, -3.0), hi=(70.4, 40.0, 1.0), duplicate_frac=0.0, quantize=None):
    g = torch.Generator().manual_seed(seed)
    lo_t, hi_t = torch.tensor(lo), torch.tensor(hi)
    xyz = lo_t + (hi_t - lo_t) * torch.rand(b, n, 3, generator=g)
    if quantize:  # snap to a grid: many exact distance ties
        xyz = torch.round(xyz / quantize) * quantize
    nd = int(n * duplicate_frac)
    if nd > 0:
        for s in range(b):
            src = torch.randint(0, n - nd, (nd,), generator=g)
            xyz[s, n - nd:] = xyz[s, src]
    return xyz.contiguous()


def random_boxes(seed, n, extent=(40.0, 40.0, 2.0), heading=True):
    """KITTI-like boxes [x,y,z,dx,dy,dz,heading]: class mean sizes x exp(N(0,0.1)), headings U(-pi,pi)."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]])
    ctr = torch.rand(n, 3, generator=g) * torch.tensor(extent)
    cls = torch.randint(0, 3, (n,), generator=g)
    size = mean[cls] * torch.exp(0.1 * torch.randn(n, 3, generator=g))
    ang = (torch.rand(n, 1, generator=g) * 2 - 1) * math.pi if heading else torch.zeros(n, 1)
    return torch.cat([ctr, size, ang], dim=1).contiguous()


def adversarial_boxes(seed, n):
    """Duplicates, shared centres, axis-aligned and nested boxes: the degenerate polygon cases."""
    b = random_boxes(seed, n, extent=(12.0, 12.0, 1.0))
    g = torch.Generator().manual_seed(seed + 1)
    q = n // 8
    b[q:2 * q] = b[:q]                                   # exact duplicates
    b[2 * q:3 * q, :2] = b[:q, :2]                       # same centre, different size / heading
    b[3 * q:4 * q, 6] = 0.0                              # axis-aligned
    b[4 * q:5 * q, 6] = math.pi / 2                      # quarter turn
    b[5 * q:6 * q] = b[:q]
    b[5 * q:6 * q, 3:5] *= 0.5                           # nested inside box i
    b[6 * q:7 * q] = b[:q]
    b[6 * q:7 * q, 0] += 1e-4 * torch.randn(q, generator=g)  # almost-duplicates
    return b.contiguous()


def boundary_boxes(seed, n):
    """Pairs of boxes whose separation straddles the far-pair early-out of csrc/nms.cu (axis-aligned squares around the
    circumscribed circles, 5 cm of slack): corner-to-corner near misses and grazing contacts at every heading."""
    g = torch.Generator().manual_seed(seed)
    b = random_boxes(seed, n, extent=(400.0, 400.0, 1.0))
    half = n // 2
    a, c = b[:half], b[half:2 * half]
    ra = 0.5 * torch.sqrt(a[:, 3] ** 2 + a[:, 4] ** 2)
    rc = 0.5 * torch.sqrt(c[:, 3] ** 2 + c[:, 4] ** 2)
    reach = (ra + rc) * 1.0001 + 0.05
    frac = 1.0 + (torch.rand(half, generator=g) - 0.5) * 0.02           # within 1 % of the cut, both sides
    frac[::5] = 0.7 + 0.3 * torch.rand(frac[::5].shape, generator=g)   # some clearly inside
    sign = torch.where(torch.rand(half, 2, generator=g) < 0.5, -1.0, 1.0)
    mode = torch.randint(0, 3, (half,), generator=g)                   # offset along x, along y, or along both
    off = torch.zeros(half, 2)
    off[:, 0] = torch.where(mode != 1, reach * frac, torch.rand(half, generator=g) * reach)
    off[:, 1] = torch.where(mode != 0, reach * frac, torch.rand(half, generator=g) * reach)
    c[:, :2] = a[:, :2] + sign * off
    return b.contiguous()


def seeded_head_model(cfg, ops=None, nms_utils=None, **kw):
    """The detector whose HEAD parameters the head golden was made with: model under manual_seed(0), then the head
    re-initialised under manual_seed(1) in the reference's construction order, BatchNorm statistics randomised from
    generator seed 3 (tests/golden/make_head_golden.py checks that the reference head built the same way is identical)."""
    from oracle import torch_ops
    from pdanet_b200.iassd import build_model
    from pdanet_b200.iassd_head import IASSD_Head
    torch.manual_seed(0)
    model = build_model(cfg, ops=ops if ops is not None else torch_ops,
                        nms_utils=nms_utils if nms_utils is not None else torch_ops.nms_utils, **kw).eval()
    torch.manual_seed(1)
    model.point_head = IASSD_Head(num_class=len(cfg.CLASS_NAMES), input_channels=model.backbone_3d.num_point_features,
                                  model_cfg=cfg.MODEL.POINT_HEAD).eval()
    model.module_list = [model.backbone_3d, model.point_head]
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for k, v in model.point_head.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif k.endswith("running_var"):
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
    return model


def head_golden_inputs(cfg, name, in_dim):
    """Seeded head inputs of the head golden: post-ReLU centre features and clustered centres inside the range."""
    g = torch.Generator().manual_seed(4)
    B = 3
    M = 256 if name == "kitti" else 1024
    lo, hi = torch.tensor(cfg.POINT_CLOUD_RANGE[:3]), torch.tensor(cfg.POINT_CLOUD_RANGE[3:])
    feats = torch.relu(torch.randn(B * M, in_dim, generator=g))
    xyz = lo + (hi - lo) * torch.rand(B * M, 3, generator=g)
    xyz[1::2] = xyz[0::2] + 0.5 * torch.randn(B * M // 2, 3, generator=g)   # neighbours: NMS has work to do
    bidx = torch.arange(B).repeat_interleave(M).float()
    centers = torch.cat([bidx[:, None], xyz], dim=1)
    return {"batch_size": B, "centers_features": feats, "centers": centers, "ctr_offsets": centers.clone(),
            "centers_origin": centers.clone(), "sa_ins_preds": [], "sample_list_id": []}

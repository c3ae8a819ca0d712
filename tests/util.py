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

"""Box coder used by the PDA-SSD head: residual centre/size + 12-bin orientation.
Decode side of the reference's PointResidual_BinOri_Coder (pcdet/utils/box_coder_utils.py:224-319);
the encode side is training-only and out of scope."""
from __future__ import annotations

import math

import torch


class PointResidual_BinOri_Coder:
    def __init__(self, code_size=8, use_mean_size=True, **kwargs):
        self.bin_size = kwargs.get("bin_size", 12)  # 'angle_bin_num' is ignored by the reference too (:227)
        self.code_size = 6 + 2 * self.bin_size
        self.bin_inter = 2 * math.pi / self.bin_size
        self.use_mean_size = use_mean_size
        if self.use_mean_size:
            self.mean_size = torch.tensor(kwargs["mean_size"], dtype=torch.float32)
            assert self.mean_size.min() > 0

    def decode_torch(self, box_encodings, points, pred_classes=None):
        """box_encodings (N, 6 + 2*bins), points (N,3), pred_classes (N) in [1, num_class] -> boxes (N,7)."""
        xt, yt, zt, dxt, dyt, dzt = torch.split(box_encodings[..., :6], 1, dim=-1)
        xa, ya, za = torch.split(points, 1, dim=-1)
        if self.use_mean_size:
            if self.mean_size.device != box_encodings.device:
                self.mean_size = self.mean_size.to(box_encodings.device)
            anchor = self.mean_size[pred_classes - 1]
            dxa, dya, dza = torch.split(anchor, 1, dim=-1)
            diagonal = torch.sqrt(dxa ** 2 + dya ** 2)
            xg = xt * diagonal + xa
            yg = yt * diagonal + ya
            zg = zt * dza + za
            dxg = torch.exp(dxt) * dxa
            dyg = torch.exp(dyt) * dya
            dzg = torch.exp(dzt) * dza
        else:
            xg, yg, zg = xt + xa, yt + ya, zt + za
            dxg, dyg, dzg = torch.split(torch.exp(box_encodings[..., 3:6]), 1, dim=-1)
        bin_scores = box_encodings[..., 6:6 + self.bin_size]
        bin_res = box_encodings[..., 6 + self.bin_size:]
        bin_id = torch.max(bin_scores, dim=-1)[1]
        res = torch.gather(bin_res, -1, bin_id.unsqueeze(-1)).squeeze(-1)
        rg = bin_id.float() * self.bin_inter - math.pi + self.bin_inter / 2
        rg = rg + res * (self.bin_inter / 2)
        return torch.cat([xg, yg, zg, dxg, dyg, dzg, rg.unsqueeze(-1)], dim=-1)

"""Seeded synthetic 16-bit-normalised star-field pairs for benchmarks and tests.

Mirrors the value distribution the reference trains on: misc/Dataset_step4_normalization.py:159-172 (log1p stretch,
clip, uint16) and dataset/astronomical_dataset_swin.py:34-39 (/65535 -> float32, shape (1,H,W)).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def synthetic_pairs(n: int, seed: int = 1234, hr_size: int = 512, scale: int = 4):
    """Returns (lr, hr) float32 in [0,1], shapes (n,1,hr/scale,hr/scale) and (n,1,hr,hr), both quantised to
    uint16 levels.  HR = background + Gaussian stars + smooth nebula, log1p-stretched; LR = box-mean / PSF / noise."""
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(hr_size, dtype=torch.float32), torch.arange(hr_size, dtype=torch.float32),
                            indexing="ij")
    hrs = []
    for _ in range(n):
        img = 0.08 + 0.01 * torch.randn(hr_size, hr_size, generator=g)
        nstars = int(torch.randint(20, 61, (1,), generator=g))
        for _s in range(nstars):
            cy, cx = (torch.rand(2, generator=g) * hr_size).tolist()
            sig = 1.0 + 3.0 * float(torch.rand(1, generator=g))
            amp = math.exp(math.log(0.05) + float(torch.rand(1, generator=g)) * (math.log(1.0) - math.log(0.05)))
            img = img + amp * torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sig * sig))
        neb = torch.randn(1, 1, hr_size // 32, hr_size // 32, generator=g)
        neb = F.interpolate(neb, size=(hr_size, hr_size), mode="bicubic", align_corners=False)[0, 0]
        img = img + 0.3 * (neb - neb.min()) / (neb.max() - neb.min() + 1e-6) * 0.5
        img = torch.log1p(img.clamp_min(0)) / math.log(2.0)
        hrs.append(img.clamp(0, 1))
    hr = torch.stack(hrs)[:, None]
    hr = torch.round(hr * 65535.0) / 65535.0
    lr = F.avg_pool2d(hr, scale)
    k = torch.arange(-4, 5, dtype=torch.float32)
    gk = torch.exp(-(k ** 2) / (2 * 1.5 ** 2)); gk = gk / gk.sum()
    lr = F.conv2d(F.pad(lr, (4, 4, 4, 4), mode="reflect"), gk.reshape(1, 1, 1, 9))
    lr = F.conv2d(lr, gk.reshape(1, 1, 9, 1))
    lr = lr + 0.005 * torch.randn(lr.shape, generator=g)
    lr = torch.round(lr.clamp(0, 1) * 65535.0) / 65535.0
    return lr, hr

"""Synthetic degraded / clean image pairs for benchmarks and smoke runs (there are no datasets offline).

Mirrors the reference's degradations: Gaussian noise sigma 15/25/50 with uint8 quantisation (utils/degradation_utils.py:21-40,
utils/dataset_utils.py:195-198), plus simple rain-streak and haze models for the other two tasks of the all-in-one setting.  The
network's cost is data independent, so the distribution only matters for parity numbers.  Host-side, pure torch, seeded."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def synthetic_batch(b: int, h: int, w: int, seed: int = 1):
    """Returns (degraded, clean) fp32 NCHW in [0,1]; degradation cycles noise15/25/50/rain/haze."""
    g = torch.Generator().manual_seed(seed)
    clean = torch.rand(b, 3, h, w, generator=g)
    clean = F.avg_pool2d(F.pad(clean, (1, 1, 1, 1), mode="reflect"), 3, stride=1)   # mild low-pass
    out = torch.empty_like(clean)
    for i in range(b):
        kind = i % 5
        c = clean[i]
        if kind < 3:
            sigma = (15.0, 25.0, 50.0)[kind]
            n = torch.randn(c.shape, generator=g)
            out[i] = torch.clamp(c * 255.0 + sigma * n, 0, 255).floor() / 255.0       # uint8 quantisation
        elif kind == 3:
            mask = (torch.rand(1, h, w, generator=g) > 0.97).float()
            streak = F.max_pool2d(mask[None], (7, 1), stride=1, padding=(3, 0))[0]
            inten = 0.5 + 0.5 * torch.rand(1, generator=g)
            out[i] = torch.clamp(c + streak * inten, 0, 1)
        else:
            t = 0.3 + 0.6 * torch.rand(1, generator=g)
            a = 0.7 + 0.3 * torch.rand(1, generator=g)
            out[i] = c * t + a * (1 - t)
    return out, clean


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """PSNR with data_range = 1 on clipped tensors."""
    mse = (a.clamp(0, 1).double() - b.clamp(0, 1).double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)

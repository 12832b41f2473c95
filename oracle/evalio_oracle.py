"""CPU oracle for the evaluation I/O helpers.  TEST INFRASTRUCTURE ONLY.

The reference scores images with scikit-image (`utils/val_utils.py:3,61-62`: `peak_signal_noise_ratio`, `structural_similarity`),
a third-party dependency that is NOT under /root/reference and not installed in this image (`env.yml` pins scikit-image 0.19.3).
Per the rules for absent dependencies this file restates the published algorithm of those two functions for the call made there
(`data_range=1, channel_axis=2`, everything else default): **parity for SSIM/PSNR is pinned to the restated formula only**, not to
skimage outputs.  `pad_to_64` and the noise rule restate reference lines directly (test.py:100-105, dataset_utils.py:195-198).
"""
from __future__ import annotations

import numpy as np
import torch


def pad_to_64(x: torch.Tensor):
    """test.py:98-105, verbatim arithmetic."""
    _, _, H, W = x.shape
    h_pad = (H // 64 + 1) * 64 - H
    w_pad = (W // 64 + 1) * 64 - W
    x = torch.cat([x, torch.flip(x, [2])], 2)[:, :, :H + h_pad, :]
    x = torch.cat([x, torch.flip(x, [3])], 3)[:, :, :, :W + w_pad]
    return x, H, W


def psnr(true: np.ndarray, test: np.ndarray, data_range: float = 1.0) -> float:
    """skimage.metrics.peak_signal_noise_ratio: 10 log10(R^2 / mean((true - test)^2)), mean in float64."""
    err = np.mean((true.astype(np.float64) - test.astype(np.float64)) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(10.0 * np.log10(data_range ** 2 / err))


def _box7(a: np.ndarray) -> np.ndarray:
    """mean over the 7x7 window centred on every pixel whose window lies inside the image ('valid' part of uniform_filter)."""
    c = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), axis=0), axis=1)
    s = c[7:, 7:] - c[:-7, 7:] - c[7:, :-7] + c[:-7, :-7]
    return s / 49.0


def ssim_channel(x: np.ndarray, y: np.ndarray, data_range: float = 1.0) -> float:
    """skimage.metrics.structural_similarity for one 2-D channel with the defaults the reference uses: win_size = 7, uniform
    filter, use_sample_covariance = True (cov_norm = 49/48), K1 = 0.01, K2 = 0.03; S is averaged over the image cropped by
    (win_size - 1) / 2 = 3 pixels, i.e. exactly the centres whose window is inside the image (so the filter's border mode is moot)."""
    x, y = x.astype(np.float64), y.astype(np.float64)
    ux, uy = _box7(x), _box7(y)
    uxx, uyy, uxy = _box7(x * x), _box7(y * y), _box7(x * y)
    cov = 49.0 / 48.0
    vx, vy, vxy = cov * (uxx - ux * ux), cov * (uyy - uy * uy), cov * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    return float(S.mean(dtype=np.float64))


def compute_psnr_ssim(recoverd: torch.Tensor, clean: torch.Tensor):
    """utils/val_utils.py:50-66."""
    assert recoverd.shape == clean.shape
    r = np.clip(recoverd.detach().cpu().numpy(), 0, 1)
    c = np.clip(clean.detach().cpu().numpy(), 0, 1)
    ps, ss = 0.0, 0.0
    for i in range(r.shape[0]):
        ps += psnr(c[i], r[i])
        ss += float(np.mean([ssim_channel(c[i, ch], r[i, ch]) for ch in range(r.shape[1])]))     # channel_axis: mean over channels
    return ps / r.shape[0], ss / r.shape[0], r.shape[0]


def add_gaussian_noise(clean255: np.ndarray, sigma: float, rng: np.random.Generator) -> np.ndarray:
    """dataset_utils.py:195-198 followed by ToTensor: uint8 truncation, then / 255."""
    noise = rng.standard_normal(clean255.shape)
    return np.clip(clean255 + noise * sigma, 0, 255).astype(np.uint8).astype(np.float32) / 255.0

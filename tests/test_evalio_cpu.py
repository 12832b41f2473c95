"""CPU: the evaluation-I/O oracle (oracle/evalio_oracle.py) against independent brute-force evaluations of the published formulas
and the reference's own padding lines; the product helpers refuse CPU tensors."""
import numpy as np
import pytest
import torch

from oracle import evalio_oracle as EO
from promptir_b200 import evalio


def test_ssim_restatement_against_brute_force():
    rng = np.random.default_rng(0)
    x, y = rng.random((19, 23)), rng.random((19, 23))
    vals = []
    for cy in range(3, 16):
        for cx in range(3, 20):
            a, b = x[cy - 3:cy + 4, cx - 3:cx + 4].ravel(), y[cy - 3:cy + 4, cx - 3:cx + 4].ravel()
            ux, uy = a.mean(), b.mean()
            vx, vy = a.var(ddof=1), b.var(ddof=1)                      # sample covariance (N - 1)
            vxy = ((a - ux) * (b - uy)).sum() / 48.0
            vals.append(((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4)))
    assert abs(EO.ssim_channel(x, y) - np.mean(vals)) < 1e-12
    assert abs(EO.ssim_channel(x, x) - 1.0) < 1e-12
    assert abs(EO.psnr(x, y) - 10 * np.log10(1.0 / np.mean((x - y) ** 2))) < 1e-12


def test_pad_rule():
    x = torch.arange(2 * 3 * 70 * 64, dtype=torch.float32).view(2, 3, 70, 64)
    p, h, w = EO.pad_to_64(x)
    assert p.shape == (2, 3, 128, 128) and (h, w) == (70, 64)              # an aligned side still gains a full 64 (test.py:100-101)
    assert torch.equal(p[..., :70, :64], x) and torch.equal(p[..., 70:128, :64], x.flip(2)[..., :58, :])
    assert torch.equal(p[..., :70, 64:], x.flip(3))


def test_no_cpu_path():
    with pytest.raises(RuntimeError):
        evalio.compute_psnr_ssim(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16))
    with pytest.raises(RuntimeError):
        evalio.pad_to_64(torch.rand(1, 3, 70, 70))

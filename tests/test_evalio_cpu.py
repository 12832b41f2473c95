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


def _ssim_the_way_skimage_computes_it(im1, im2, dtype):
    """The body of skimage.metrics.structural_similarity (0.19.x, the version env.yml pins) for one channel with the arguments the
    reference passes, written with the SciPy primitive skimage itself calls: `scipy.ndimage.uniform_filter(size=7)` (border mode
    'reflect'), sample covariance, crop by (win_size - 1) // 2, float64 mean.  skimage keeps float32 inputs in float32
    (`_supported_float_type`); the oracle works in float64 -- both are evaluated here."""
    from scipy.ndimage import uniform_filter
    im1, im2 = im1.astype(dtype), im2.astype(dtype)
    NP = 7 ** 2
    cov_norm = NP / (NP - 1)
    ux, uy = uniform_filter(im1, size=7), uniform_filter(im2, size=7)
    uxx, uyy, uxy = uniform_filter(im1 * im1, size=7), uniform_filter(im2 * im2, size=7), uniform_filter(im1 * im2, size=7)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * 1.0) ** 2, (0.03 * 1.0) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    return float(S[3:-3, 3:-3].mean(dtype=np.float64))


def test_ssim_restatement_against_the_scipy_filter_skimage_uses():
    """Second, independent pin of the SSIM restatement (skimage itself is absent here): the same quantity through
    scipy.ndimage.uniform_filter, ragged and minimum (7x7: a single window) sizes, float64 exactly and float32 to rounding."""
    rng = np.random.default_rng(1)
    for h, w in ((7, 7), (7, 31), (19, 23), (64, 48), (321, 481)):
        x = rng.random((h, w)).astype(np.float32)
        y = np.clip(x + 0.1 * rng.standard_normal((h, w)), 0, 1).astype(np.float32)
        got = EO.ssim_channel(x, y)
        assert abs(got - _ssim_the_way_skimage_computes_it(x, y, np.float64)) < 1e-12, (h, w)
        assert abs(got - _ssim_the_way_skimage_computes_it(x, y, np.float32)) < 5e-6, (h, w)


def test_ssim_psnr_known_answers():
    """Closed forms of the published definitions: constant images have zero variance, so SSIM reduces to the luminance term
    (2ab + C1) / (a^2 + b^2 + C1); PSNR of a constant offset d is -20 log10(d); identical images give SSIM 1 and PSNR inf."""
    a, b = 0.25, 0.75
    x, y = np.full((16, 20), a), np.full((16, 20), b)
    assert abs(EO.ssim_channel(x, y) - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-12
    assert abs(EO.psnr(x, y) - (-20 * np.log10(b - a))) < 1e-12
    assert EO.ssim_channel(x, x) == 1.0 and EO.psnr(x, x) == float("inf")
    # the batch helper: mean over channels, then over images (val_utils.py:58-64)
    r, c = torch.full((2, 3, 16, 20), a), torch.full((2, 3, 16, 20), b)
    ps, ss, n = EO.compute_psnr_ssim(r, c)
    assert n == 2 and abs(ps + 20 * np.log10(0.5)) < 1e-6 and abs(ss - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-9


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

"""GPU parity of the on-device evaluation I/O (promptir_b200/evalio.py) against oracle/evalio_oracle.py."""
import pytest
import torch

from oracle import evalio_oracle as EO
from promptir_b200 import evalio

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape", [(2, 3, 70, 64), (1, 3, 100, 37 + 64), (1, 1, 64, 64)])
def test_pad_to_64(shape):
    x = torch.rand(*shape)
    got, h, w = evalio.pad_to_64(x.to(DEV))
    ref, h2, w2 = EO.pad_to_64(x)
    assert (h, w) == (h2, w2) and torch.equal(got.cpu(), ref)                   # a permutation: bit exact


@pytest.mark.parametrize("shape", [(2, 3, 64, 96), (3, 3, 37, 53), (1, 1, 7, 7), (1, 3, 321, 481)])
def test_psnr_ssim(shape):
    torch.manual_seed(shape[2])
    clean = torch.rand(*shape)
    rec = (clean + 0.1 * torch.randn(*shape)) * 1.05 - 0.02                     # leaves [0, 1]: exercises the clip
    p, s, n = evalio.compute_psnr_ssim(rec.to(DEV), clean.to(DEV))
    pr, sr, nr = EO.compute_psnr_ssim(rec, clean)
    assert n == nr and abs(p - pr) <= 1e-4 and abs(s - sr) <= 2e-5, (p, pr, s, sr)
    per = evalio.psnr_ssim_per_image(clean.to(DEV), clean.to(DEV))
    assert torch.isinf(per[:, 0]).all() and (per[:, 1] - 1).abs().max().item() < 1e-6       # identical images: skimage gives inf, 1


def test_add_gaussian_noise_statistics():
    clean = torch.full((4, 3, 128, 128), 128.0, device=DEV)
    for sigma in (15, 25, 50):
        out = evalio.add_gaussian_noise(clean, sigma, seed=sigma)
        v = out * 255
        assert (v - v.round()).abs().max().item() < 1e-3 and out.min().item() >= 0 and out.max().item() <= 1
        assert abs(v.mean().item() - 127.5) < 0.6 and abs(v.std().item() - sigma) < 0.02 * sigma + 0.3      # floor() shifts the mean by -0.5
        assert torch.equal(out, evalio.add_gaussian_noise(clean, sigma, seed=sigma))                         # counter-based: reproducible
        assert not torch.equal(out, evalio.add_gaussian_noise(clean, sigma, seed=sigma + 1))
    dark = evalio.add_gaussian_noise(torch.zeros(1, 3, 64, 64, device=DEV), 25, seed=1)
    assert 0.4 < (dark == 0).float().mean().item() < 0.6                                                      # clip at 0


def test_eval_loop_matches_host_scoring():
    """The on-device evaluation step (noise -> pad -> forward -> crop -> score) against the same images scored by the oracle on the host."""
    from promptir_b200 import PromptIR, synth
    torch.manual_seed(0)
    net = PromptIR(decoder=True).eval().to(DEV)
    net.compute_dtype = torch.float16
    _, clean = synth.synthetic_batch(2, 70, 100, seed=5)
    clean = clean.to(DEV)
    with torch.no_grad():
        degrad = evalio.add_gaussian_noise(clean * 255.0, 25.0, seed=3)
        padded, h, w = evalio.pad_to_64(degrad)
        assert padded.shape[-2:] == (128, 128) and (h, w) == (70, 100)
        restored = net(padded)[:, :, :h, :w]
        p, s, n = evalio.compute_psnr_ssim(restored, clean)
    pr, sr, nr = EO.compute_psnr_ssim(restored.cpu(), clean.cpu())
    assert n == nr == 2 and abs(p - pr) <= 1e-4 and abs(s - sr) <= 2e-5
    pd, sd_, _ = evalio.compute_psnr_ssim(degrad, clean)
    assert 19.5 < pd < 21.5                                   # sigma 25 on [0, 255] is ~20.2 dB before clipping effects

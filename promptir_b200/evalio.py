"""On-device evaluation I/O around the restoration forward.

Mirrors the reference's helpers with the same names, argument meaning and return values, but everything stays on the GPU:
  pad_to_64(x)                      test.py:98-105   (flip-concat padding to the next multiple of 64 -- a FULL extra 64 when already aligned)
  compute_psnr_ssim(recoverd, clean)  utils/val_utils.py:50-66  -> (mean psnr, mean ssim, N), skimage definitions
  add_gaussian_noise(clean255, sigma) utils/dataset_utils.py:195-198 / degradation_utils.py:21-26
No CPU fallback: CPU tensors raise."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"promptir_b200.evalio.{what}: tensor is on the CPU and there is no CPU path")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def pad_to_64(x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """-> (padded, H_old, W_old).  padded[..., :H_old, :W_old] is x; the rest mirrors it (test.py:100-105)."""
    _need_cuda(x, "pad_to_64")
    x = x.float().contiguous()
    B, C, H, W = x.shape
    Hp, Wp = (H // 64 + 1) * 64, (W // 64 + 1) * 64
    if Hp > 2 * H or Wp > 2 * W:
        raise RuntimeError(f"pad_to_64: cannot mirror {H}x{W} up to {Hp}x{Wp} (the reference's flip-concat has the same limit)")
    out = torch.empty(B, C, Hp, Wp, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().pir_mirror_pad(x.data_ptr(), out.data_ptr(), B * C, H, W, Hp, Wp, _stream(x)), "pir_mirror_pad")
    _lib.launch_count += 1
    return out, H, W


def psnr_ssim_per_image(recoverd: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    """-> float64 [B, 2] on the device: (psnr, ssim) of every image."""
    _need_cuda(recoverd, "compute_psnr_ssim")
    assert recoverd.shape == clean.shape and recoverd.dim() == 4
    a, b = recoverd.detach().float().contiguous(), clean.detach().float().contiguous()
    B, C, H, W = a.shape
    lib = _lib.load()
    ws = torch.empty(int(lib.pir_psnr_ssim_ws_bytes(B, C, H, W)), dtype=torch.uint8, device=a.device)
    out = torch.empty(B, 2, dtype=torch.float64, device=a.device)
    _lib.check(lib.pir_psnr_ssim(a.data_ptr(), b.data_ptr(), B, C, H, W, ws.data_ptr(), out.data_ptr(), _stream(a)), "pir_psnr_ssim")
    _lib.launch_count += 2
    return out


def compute_psnr_ssim(recoverd: torch.Tensor, clean: torch.Tensor) -> Tuple[float, float, int]:
    """utils/val_utils.py:50-66: batch means of skimage's PSNR and SSIM on clip(., 0, 1) images, and the batch size."""
    per = psnr_ssim_per_image(recoverd, clean)
    m = per.mean(dim=0).tolist()                           # the only device -> host transfer: two doubles
    return m[0], m[1], recoverd.shape[0]


def add_gaussian_noise(clean255: torch.Tensor, sigma: float, seed: int = 0) -> torch.Tensor:
    """clean255: any-shape float tensor of 0..255 pixel values -> noisy image in [0, 1] quantised to 1/255 steps
    (`np.clip(clean + noise * sigma, 0, 255).astype(np.uint8)` then ToTensor's / 255)."""
    _need_cuda(clean255, "add_gaussian_noise")
    c = clean255.float().contiguous()
    out = torch.empty_like(c)
    _lib.check(_lib.load().pir_add_noise(c.data_ptr(), out.data_ptr(), c.numel(), float(sigma), int(seed) & (2 ** 64 - 1), _stream(c)), "pir_add_noise")
    _lib.launch_count += 1
    return out

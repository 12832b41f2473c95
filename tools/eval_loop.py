#!/usr/bin/env python
"""The reference's denoising evaluation loop (test.py:80-118: add noise -> pad to 64 -> net -> crop -> PSNR/SSIM -> average)
with every step on the GPU: promptir_b200.evalio supplies the noise synthesis, the flip-concat padding and the skimage-definition
PSNR / SSIM, so nothing but two doubles per image crosses to the host.  Synthetic clean images (no datasets offline).

    python tools/eval_loop.py [--images 8] [--height 321] [--width 481] [--sigma 25] [--dtype bf16]
Prints one JSON line: mean PSNR / SSIM of the degraded input and of the (random-init) restoration, images per second, and the
share of the loop spent outside the network forward."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


class AverageMeter:                                   # utils/val_utils.py:7-25
    def __init__(self):
        self.sum, self.count = 0.0, 0

    def update(self, val, n=1):
        self.sum += val * n
        self.count += n

    @property
    def avg(self):
        return self.sum / max(self.count, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--height", type=int, default=321)
    ap.add_argument("--width", type=int, default=481)
    ap.add_argument("--sigma", type=float, default=25.0)
    ap.add_argument("--dtype", default="bf16")
    args = ap.parse_args()
    from promptir_b200 import PromptIR, evalio, synth
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = PromptIR(decoder=True).eval().to(dev)
    net.compute_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    _, clean_all = synth.synthetic_batch(args.images, args.height, args.width, seed=5)
    clean_all = clean_all.to(dev)
    psnr_in, ssim_in, psnr_out, ssim_out = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
    t_net = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(args.images):                                   # batch_size = 1, as test.py:90
            clean = clean_all[i:i + 1]
            degrad = evalio.add_gaussian_noise(clean * 255.0, args.sigma, seed=i)          # dataset_utils.py:195-198
            padded, h_old, w_old = evalio.pad_to_64(degrad)                                # test.py:98-105
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            restored = net(padded)[:, :, :h_old, :w_old]                                   # test.py:109-110
            torch.cuda.synchronize()
            t_net += time.perf_counter() - t1
            p, s, n = evalio.compute_psnr_ssim(restored, clean)                            # test.py:111
            psnr_out.update(p, n)
            ssim_out.update(s, n)
            p, s, n = evalio.compute_psnr_ssim(degrad, clean)
            psnr_in.update(p, n)
            ssim_in.update(s, n)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(json.dumps({"images": args.images, "size": [args.height, args.width], "sigma": args.sigma,
                      "degraded": {"psnr": psnr_in.avg, "ssim": ssim_in.avg}, "restored_random_init": {"psnr": psnr_out.avg, "ssim": ssim_out.avg},
                      "images_per_sec": args.images / total, "seconds": total, "share_outside_forward": 1.0 - t_net / total}))


if __name__ == "__main__":
    main()

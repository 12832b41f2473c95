"""Time the dense 3x3 convolutions of the cfg2 forward one by one (diagnostic).  python tools/time_conv3.py [fp16|bf16]
Honours PIR_CONV3=0 (old one-tile-per-CTA kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptir_b200 import ops, packing  # noqa: E402
from promptir_b200._lib import OUT_FINAL_NCHW32, OUT_NHWC16, OUT_SHUFFLE16, OUT_UNSHUFFLE16  # noqa: E402

CASES = [("down1_2", 256, 48, 24, OUT_UNSHUFFLE16), ("down2_3", 128, 96, 48, OUT_UNSHUFFLE16), ("up2_1", 128, 96, 192, OUT_SHUFFLE16),
         ("prompt1", 128, 64, 64, OUT_NHWC16), ("prompt2", 64, 128, 128, OUT_NHWC16), ("output", 256, 96, 3, OUT_FINAL_NCHW32)]


def main():
    dt = torch.float16 if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else torch.bfloat16
    B = 16
    s = torch.cuda.current_stream().cuda_stream
    for name, side, cin, cout, mode in CASES:
        a = torch.randn(B, side, side, cin, device="cuda").to(dt)
        w16 = packing.pack_conv3x3(torch.randn(cout, cin, 3, 3, device="cuda") / (9 * cin) ** 0.5, dt)
        img = None
        if mode == OUT_NHWC16:
            out = torch.zeros(B, side, side, cout, device="cuda", dtype=dt)
        elif mode == OUT_UNSHUFFLE16:
            out = torch.zeros(B, side // 2, side // 2, 4 * cout, device="cuda", dtype=dt)
        elif mode == OUT_SHUFFLE16:
            out = torch.zeros(B, 2 * side, 2 * side, cout // 4, device="cuda", dtype=dt)
        else:
            out = torch.zeros(B, cout, side, side, device="cuda")
            img = torch.rand(B, cout, side, side, device="cuda")
        launch = ops.gemm(a, w16, out, n=cout, taps=9, out_mode=mode, img=img)
        for _ in range(3):
            launch(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            launch(s)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        by = a.numel() * 2 + out.numel() * out.element_size() + (0 if img is None else img.numel() * 4)
        print(f"conv3x3 {name:8s} {side}x{side} {cin}->{cout}: {us:7.1f} us  {by / us / 1e3:7.1f} GB/s  (PIR_CONV3={os.environ.get('PIR_CONV3', '1')})")


if __name__ == "__main__":
    main()

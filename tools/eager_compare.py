#!/usr/bin/env python
"""Same-box comparison asked for in BASELINE.md §3: the reference network as PyTorch eager ops (cuDNN / cuBLAS) on the B200 next to
this library, BASELINE config 2 (batch 16 of 256x256).  /root/reference does not exist on the GPU box, so the eager side is the
oracle port -- a functional torch restatement of net/model.py pinned to the real reference (checker-side code; nothing here is on
the product path).  Prints one JSON line.

    python tools/eager_compare.py [--batch 16] [--side 256] [--steps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def timed(fn, steps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--side", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--train", action="store_true", help="forward + L1 loss + backward (BASELINE config 4: --batch 32 --side 128) instead of the forward")
    args = ap.parse_args()
    from oracle import promptir_oracle as O
    from promptir_b200 import PromptIR, synth
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = PromptIR(decoder=True).eval().to(dev)
    x, _ = synth.synthetic_batch(args.batch, args.side, args.side, seed=1)
    x = x.to(dev)
    mp = args.batch * args.side * args.side / 1e6
    if args.train:
        import torch.nn.functional as F
        _, y = synth.synthetic_batch(args.batch, args.side, args.side, seed=1)
        y = y.to(dev)
        res = {"workload": f"batch {args.batch} of {args.side}x{args.side}, forward + L1 loss + backward, one B200"}
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True

        def eager_step(sd, xin, yin):
            for v in sd.values():
                v.grad = None
            F.l1_loss(O.promptir_forward(sd, xin).float(), yin).backward()
        sd32 = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        ms = timed(lambda: eager_step(sd32, x, y), args.steps)
        res["eager_autograd_fp32_tf32"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
        sd16 = {k: v.detach().to(torch.bfloat16).requires_grad_(True) for k, v in net.state_dict().items()}
        xb = x.to(torch.bfloat16)
        ms = timed(lambda: eager_step(sd16, xb, y), args.steps)
        res["eager_autograd_bf16"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
        del sd32, sd16
        torch.cuda.empty_cache()
        net.train()
        net.compute_dtype = torch.bfloat16

        def ours():
            net.zero_grad(set_to_none=True)
            F.l1_loss(net(x), y).backward()
        ms = timed(ours, args.steps)
        res["promptir_b200_bf16"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
        res["speedup_vs_best_eager"] = min(res["eager_autograd_fp32_tf32"]["ms"], res["eager_autograd_bf16"]["ms"]) / ms
        print(json.dumps(res))
        return
    res = {"workload": f"batch {args.batch} of {args.side}x{args.side}, forward, one B200"}
    with torch.no_grad():
        sd32 = {k: v.detach() for k, v in net.state_dict().items()}
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        ms = timed(lambda: O.promptir_forward(sd32, x), args.steps)
        res["eager_fp32_tf32"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
        sd16 = {k: v.to(torch.bfloat16) for k, v in sd32.items()}
        xb = x.to(torch.bfloat16)
        ms = timed(lambda: O.promptir_forward(sd16, xb), args.steps)
        res["eager_bf16"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
        sdcl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd16.items()}
        xcl = xb.contiguous(memory_format=torch.channels_last)
        ms = timed(lambda: O.promptir_forward(sdcl, xcl), args.steps)
        res["eager_bf16_channels_last"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
        net.compute_dtype = torch.bfloat16
        ms = timed(lambda: net(x), args.steps)
        res["promptir_b200_bf16"] = {"ms": ms, "MP_per_s": mp / ms * 1e3}
    best = min(res[k]["ms"] for k in ("eager_fp32_tf32", "eager_bf16", "eager_bf16_channels_last"))
    res["speedup_vs_best_eager"] = best / res["promptir_b200_bf16"]["ms"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()

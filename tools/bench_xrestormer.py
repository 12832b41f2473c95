#!/usr/bin/env python
"""PromptXRestormer forward benchmark (BASELINE.json configs[4]: bf16 inference, batch 8 at 512x512 over 8 B200 = 1 image per GPU).
One JSON line on stdout (rank 0): MP/s with inputs resident (CUDA-graph replay, CUDA events, max over ranks), the per-kernel
table with achieved GB/s / TFLOP/s, and parity of image 0 against the fp32 CPU oracle on a 128x128 crop.

    python tools/bench_xrestormer.py [--batch 1] [--side 512] [--steps 10] [--warmup 3] [--dtype bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_xrestormer.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--side", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    args = ap.parse_args()
    real_stdout = os.dup(1)          # libraries (NCCL banner) print to fd 1: keep stdout for the one JSON line
    os.dup2(2, 1)
    from promptir_b200 import PromptXRestormer, synth
    from promptir_b200.xengine import x_op_cost

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = PromptXRestormer().eval().to(dev)
    net.compute_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    B, S = args.batch, args.side
    x, _ = synth.synthetic_batch(B, S, S, seed=1 + rank)
    x = x.to(dev)
    eng = net.engine_for(B, S, S, dev)
    eng.img_in.copy_(x)
    for _ in range(args.warmup):
        eng.replay(True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.replay(True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        stream = torch.cuda.current_stream().cuda_stream
        for rep in range(2):
            evs = []
            for r in eng.ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); r["launch"](stream); b.record()
                evs.append((r, a, b))
            torch.cuda.synchronize()
        kernels, tot = {}, 0.0
        for r, a, b in evs:
            t = a.elapsed_time(b)
            tot += t
            by, fl = x_op_cost(r)
            k = kernels.setdefault(r.get("tag") or r["kind"], {"launches": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
            k["launches"] += 1; k["ms"] += t; k["bytes"] += by; k["flops"] += fl
        for k in kernels.values():
            k["GBps"] = round(k.pop("bytes") / k["ms"] / 1e6, 1)
            k["TFLOPs"] = round(k.pop("flops") / k["ms"] / 1e9, 2)
            k["ms"] = round(k["ms"], 3)
            k["share"] = round(k["ms"] / tot, 4)
        kernels = dict(sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]))
        # parity on a crop the CPU oracle finishes in seconds (checker leg: the only user of oracle/)
        from oracle import xrestormer_oracle as XO
        xc = x[:1, :, :128, :128].contiguous()
        with torch.no_grad():
            yc = net(xc).cpu()
            ref = XO.xrestormer_forward({k: v.detach().cpu() for k, v in net.state_dict().items()}, xc.cpu())
        os.write(real_stdout, (json.dumps({"metric": "prompt_xrestormer_fwd_megapixels_per_sec", "value": world * B * S * S / 1e6 / (ms / 1e3), "unit": "MP/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "dtype": args.dtype, "data": "synthetic",
                          "config": {"workload": f"PromptXRestormer (dim 48, [4,6,6,8]) inference, batch {B} of {S}x{S} per GPU "
                                                 f"(BASELINE.json configs[4]), random-init weights seed 0",
                                     "parallelism": f"images sharded over {world} GPU(s), no data-path collective"},
                          "gpu_launches": eng.kernels_per_forward(), "kernels": kernels,
                          "parity": {"max_abs_clamped": (yc.clamp(0, 1) - ref.clamp(0, 1)).abs().max().item(), "oracle": "fp32 CPU port, 128x128 crop"}}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

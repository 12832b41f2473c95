#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3]): forward + L1 loss + backward on batches of 128x128 patches, 32 per GPU,
gradients all-reduced over NCCL when launched with torchrun.  One JSON line on stdout (rank 0).

    python tools/bench_train.py [--batch 32] [--side 128] [--steps 10] [--warmup 3] [--dtype bf16|fp16] [--adamw] [--cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py ...

  engine : TrainEngine.forward -> L1 loss (torch, 3-channel image: the caller's loss as in train.py:43) -> TrainEngine.backward,
           both programs replayed as CUDA graphs; CUDA events, max over ranks.  With N ranks the flat fp32 gradient buffer is
           all-reduced with ONE NCCL call per step inside the timed region.
  api    : the same step through the drop-in module: loss = L1(net(x), y); loss.backward()   (what train.py does)
  per-kernel table: an eager pass with CUDA events around every launch, grouped by tag.
`--cpu-baseline` also times fp32 autograd of the oracle port on the host cores for ONE patch (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--side", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--adamw", action="store_true", help="include torch.optim.AdamW(fused=True).step() in the timed step")
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--model", default="promptir", choices=["promptir", "xrestormer"], help="xrestormer: the PromptXRestormer variant (side % 64 == 0)")
    ap.add_argument("--ops-file", default="", help="write the per-launch timing list (tag, kind, shape, ms) here")
    args = ap.parse_args()
    real_stdout = os.dup(1)          # libraries (NCCL banner) print to fd 1: keep stdout for the one JSON line
    os.dup2(2, 1)

    from promptir_b200 import PromptIR, _lib, synth
    from promptir_b200.train_engine import TrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    B, S = args.batch, args.side
    torch.manual_seed(0)
    if args.model == "xrestormer":
        from promptir_b200 import PromptXRestormer
        from promptir_b200.xtrain_engine import XTrainEngine as TrainEngine      # noqa: F811
        net = PromptXRestormer().to(dev).train()
    else:
        net = PromptIR(decoder=True).to(dev).train()
    net.compute_dtype = dt
    x, y = synth.synthetic_batch(B, S, S, seed=1 + rank)
    x, y = x.to(dev), y.to(dev)
    eng = TrainEngine(net, B, S, S, dev, dt)
    params = [p for n, p in net.named_parameters() if n in eng.live_params]
    for n, p in net.named_parameters():                      # zero-copy: .grad are views of the flat buffer
        p.grad = eng.grads[n] if n in eng.live_params else None
    opt = torch.optim.AdamW(params, lr=2e-4, fused=True) if args.adamw else None

    def step_engine():
        out = eng.forward(x)
        out.requires_grad_(True)
        loss = F.l1_loss(out, y)
        (d_out,) = torch.autograd.grad(loss, out)
        eng.backward(d_out)
        if world > 1:
            dist.all_reduce(eng.grad_flat)
            eng.grad_flat.mul_(1.0 / world)
        if opt is not None:
            opt.step()
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.launch_count
        e0.record()
        for _ in range(steps):
            loss = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, (_lib.launch_count - c0) // steps, loss.item()

    ms_eng, launches, loss = timed(step_engine, args.steps, args.warmup)

    # forward / backward split (graph replays alone)
    def only(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms_fwd = only(lambda: eng._run("fwd", eng.fwd_launches, True))
    ms_bwd = only(lambda: eng._run("bwd", eng.bwd_launches, True))

    # through the module API (autograd Function), single rank semantics (no all-reduce here: DDP would add it)
    del eng
    for p in net.parameters():
        p.grad = None
    torch.cuda.empty_cache()

    def step_api():
        net.zero_grad(set_to_none=True)
        loss = F.l1_loss(net(x), y)
        loss.backward()
        return loss
    ms_api, _, loss_api = timed(step_api, max(2, args.steps // 2), 2)
    eng = net._train_engine[1]

    kernels = {}
    if not args.no_kernels and rank == 0:
        stream = torch.cuda.current_stream().cuda_stream
        eng.d_out.fill_(1.0 / eng.d_out.numel())
        for rep in range(2):
            evs = []
            for r in eng.fwd_ops + eng.bwd_ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); r["launch"](stream); b.record()
                evs.append((r, a, b))
            torch.cuda.synchronize()
        tot = 0.0
        oplist = []
        for r, a, b in evs:
            t = a.elapsed_time(b)
            tot += t
            shp = next((tuple(r[k].shape) for k in ("a", "x", "out", "g", "y", "dup", "qkv") if isinstance(r.get(k), torch.Tensor)), ())
            oplist.append({"tag": r.get("tag") or r["kind"], "kind": r["kind"], "shape": shp, "ms": round(t, 4),
                           **{k: r[k] for k in ("M", "N", "P", "taps", "splits", "parts", "n") if isinstance(r.get(k), int)}})
            key = (r.get("tag") or r["kind"])
            k = kernels.setdefault(key, {"launches": 0, "ms": 0.0})
            k["launches"] += 1
            k["ms"] += t
        for k in kernels.values():
            k["ms"] = round(k["ms"], 3)
            k["share"] = round(k["ms"] / tot, 4)
        kernels = dict(sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]))
        if args.ops_file:
            with open(args.ops_file, "w") as f:
                for o in oplist:
                    f.write(json.dumps(o) + "\n")

    cpu = None
    if args.cpu_baseline and rank == 0:
        from oracle import promptir_oracle as O          # the CPU leg is the only user of oracle/
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        xc, yc = x[:1].cpu(), y[:1].cpu()
        if args.model == "xrestormer":
            from oracle import xrestormer_oracle as XO
            fwd = XO.xrestormer_forward
        else:
            fwd = O.promptir_forward
        times = []
        for i in range(3):
            t0 = time.perf_counter()
            F.l1_loss(fwd(sd, xc), yc).backward()
            times.append(time.perf_counter() - t0)
        sec = statistics.median(times[1:])
        cpu = {"value": S * S / 1e6 / sec, "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": f"1 of the {B} {S}x{S} patches per step, fp32 autograd of the oracle port, median of 2 ({sec:.2f} s each)"}

    if rank == 0:
        mp = world * B * S * S / 1e6
        os.write(real_stdout, (json.dumps({
            "metric": ("prompt_xrestormer" if args.model == "xrestormer" else "promptir") + "_train_step_megapixels_per_sec", "value": mp / (ms_eng / 1e3), "unit": "MP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_eng, "higher_is_better": True, "scaling": "weak",
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{'PromptXRestormer' if args.model == 'xrestormer' else 'PromptIR'} training step (forward + L1 loss + backward{' + AdamW' if args.adamw else ''}), "
                                   f"{S}x{S} patches, batch {B} per GPU (BASELINE.json configs[3])",
                       "parallelism": f"data parallel over {world} GPU(s), one NCCL all-reduce of the flat fp32 gradient buffer per step"},
            "images_per_sec": world * B / (ms_eng / 1e3), "ms_forward": ms_fwd, "ms_backward": ms_bwd,
            "api": {"ms_per_step": ms_api, "value": B * S * S / 1e6 / (ms_api / 1e3), "call": "loss = l1(net(x), y); loss.backward()"},
            "gpu_launches": eng.kernels_per_step(), "loss": loss, "loss_api": loss_api,
            "saved_activation_GB": eng.saved_bytes / 1e9, "cpu_baseline": cpu, "kernels": kernels}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

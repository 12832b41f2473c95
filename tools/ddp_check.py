#!/usr/bin/env python
"""2+ GPU check of the training path under data parallelism (launch with torchrun, one rank per GPU, NCCL):
  (1) torch DistributedDataParallel(find_unused_parameters=True) around the drop-in module -- what Lightning's DDP strategy does
      for train.py:307-341 -- gives every rank the same averaged gradients;
  (2) they equal the engine's own exchange: one all-reduce of the flat gradient buffer (promptir_b200/ddp.py).
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    from promptir_b200 import PromptIR, ddp, synth
    from promptir_b200.train_engine import TrainEngine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = PromptIR(decoder=True).to(dev).train()
    B, S = 4, 64
    x, y = synth.synthetic_batch(B, S, S, seed=10 + rank)
    x, y = x.to(dev), y.to(dev)

    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)
    for _ in range(2):                                   # second iteration: DDP's bucket rebuild path
        model.zero_grad(set_to_none=True)
        F.l1_loss(model(x), y).backward()
    g_ddp = torch.cat([p.grad.reshape(-1) for n, p in net.named_parameters() if p.grad is not None])
    dead = [n for n, p in net.named_parameters() if p.grad is None]
    ref = g_ddp.clone()
    dist.broadcast(ref, 0)
    same_across_ranks = bool((g_ddp == ref).all().item())

    eng = TrainEngine(net, B, S, S, dev, net.compute_dtype)
    out = eng.forward(x).requires_grad_(True)
    (d_out,) = torch.autograd.grad(F.l1_loss(out, y), out)
    eng.backward(d_out)
    ddp.allreduce_gradients(eng)
    g_flat = torch.cat([eng.grads[n].reshape(-1) for n, _ in net.named_parameters() if n in eng.live_params])
    rel = ((g_flat - g_ddp).norm() / g_ddp.norm()).item()
    ok = torch.tensor([int(same_across_ranks and rel < 1e-5 and len(dead) == 6)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "ddp_grads_identical_across_ranks": same_across_ranks, "flat_allreduce_vs_ddp_rel": rel,
                          "dead_params": len(dead), "ok": bool(ok.item())}))
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()

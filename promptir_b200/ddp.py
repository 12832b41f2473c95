"""Data-parallel gradient exchange for the training program (train.py:307-341 runs Lightning DDP: one process per GPU, bucketed
NCCL all-reduce of 142 MB of fp32 gradients hooked into autograd).

The TrainEngine writes every parameter gradient into ONE flat fp32 buffer (`grad_flat`), so the exchange is a single
`all_reduce` on that buffer -- no buckets, no per-parameter hooks, nothing to flatten -- followed by a 1/world scale (DDP averages).
NVSwitch reduces in-network (NVLS), so one 142 MB collective costs ~0.3 ms against a step of tens of ms; it is issued on the
compute stream right after the backward graph.  The module also works under `torch.nn.parallel.DistributedDataParallel(
find_unused_parameters=True)` unchanged (the reference needs that flag too: its six dead convs never get gradients).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def attach_flat_grads(module: torch.nn.Module, engine) -> None:
    """Make every live parameter's .grad a view of the engine's flat buffer (zero-copy for optimizers); dead ones get None."""
    for n, p in module.named_parameters():
        p.grad = engine.grads[n] if n in engine.live_params else None


def allreduce_gradients(engine, group: Optional[dist.ProcessGroup] = None) -> None:
    """Average `engine.grad_flat` over the data-parallel group with one collective."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    dist.all_reduce(engine.grad_flat, group=group)
    engine.grad_flat.mul_(1.0 / world)

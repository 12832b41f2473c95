"""Data-parallel gradient exchange for the training program (train.py:307-341 runs Lightning DDP: one process per GPU, bucketed
NCCL all-reduce of 142 MB of fp32 gradients hooked into autograd).

The TrainEngine writes every parameter gradient into ONE flat fp32 buffer (`grad_flat`), so the exchange is a single
`all_reduce` on that buffer -- no buckets, no per-parameter hooks, nothing to flatten -- followed by a 1/world scale (DDP averages).
NVSwitch reduces in-network (NVLS), so one 142 MB collective costs ~0.3 ms against a step of tens of ms; it is issued on the
compute stream right after the backward graph.  The module also works under `torch.nn.parallel.DistributedDataParallel(
find_unused_parameters=True)` unchanged (the reference needs that flag too: its six dead convs never get gradients).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def attach_flat_grads(module: torch.nn.Module, engine) -> None:
    """Make every live parameter's .grad a view of the engine's flat buffer (zero-copy for optimizers); dead ones get None.

    The backward program OVERWRITES the flat buffer every step: gradient accumulation over micro-batches needs the caller to sum
    copies (or use the autograd path, whose node hands clones to AccumulateGrad).  `zero_grad(set_to_none=True)` detaches the views;
    call this again afterwards (zeroing is unnecessary, the buffer is rewritten)."""
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


def plan_overlap(bwd_ops, grad_flat: torch.Tensor, segments: int = 4, grads=None) -> List[Tuple[int, int, List[Tuple[int, int]]]]:
    """Split the backward records into `segments` consecutive ranges and list, per range, the runs of `grad_flat` (element offset,
    length) whose LAST writer lies in that range -- i.e. the gradients that are final once the range has run.
    -> [(first_op, end_op, [(offset, numel), ...]), ...]; adjacent parameters are merged into one run.
    `grads` ({name: view of grad_flat}): destinations are widened to the whole parameter they lie in (a parameter is exchanged once,
    after its last writer, even when several records write parts of it)."""
    base, item, total = grad_flat.data_ptr(), grad_flat.element_size(), grad_flat.numel()
    owners = sorted(((v.data_ptr() - base) // item, v.numel()) for v in grads.values()) if grads else []
    last = {}
    for i, r in enumerate(bwd_ops):
        for k, v in r.items():
            if k.startswith("dst_") and isinstance(v, torch.Tensor):
                off, ln = (v.data_ptr() - base) // item, v.numel()
                if off < 0 or off + ln > total:
                    continue                              # not a gradient (scratch destination)
                for o, n in owners:
                    if o <= off and off + ln <= o + n:
                        off, ln = o, n
                        break
                last[(off, ln)] = i
    n = len(bwd_ops)
    bounds = [round(n * (s + 1) / segments) for s in range(segments)]
    plan, a = [], 0
    for b in bounds:
        rng = sorted(k for k, i in last.items() if a <= i < b)
        runs: List[Tuple[int, int]] = []
        for off, ln in rng:
            if runs and runs[-1][0] + runs[-1][1] == off:
                runs[-1] = (runs[-1][0], runs[-1][1] + ln)
            else:
                runs.append((off, ln))
        plan.append((a, b, runs))
        a = b
    return plan


class OverlappedReducer:
    """Data-parallel backward with the gradient exchange overlapped (train.py:339 runs DDP, whose buckets overlap the same way).

    The backward program is cut into `segments` pieces (each its own CUDA graph).  As soon as a piece has run, the runs of the
    flat gradient buffer it completed are all-reduced (average) on a side stream while the next piece computes; the compute stream
    joins the side stream at the end.  Parameters are registered in forward order and the backward walks them in reverse, so the
    completed gradients of a piece form a handful of contiguous runs.  On CPU (gloo tests) `run_range(a, b)` interprets the records
    and the reduction runs inline."""

    def __init__(self, engine, segments: int = 4, group: Optional[dist.ProcessGroup] = None, run_range=None):
        self.eng, self.group = engine, group
        self.plan = plan_overlap(engine.bwd_ops, engine.grad_flat, segments, getattr(engine, "grads", None))
        self.run_range = run_range or engine.run_bwd_range
        self.cuda = engine.grad_flat.is_cuda
        self.side = torch.cuda.Stream(engine.grad_flat.device) if self.cuda else None
        self.world = dist.get_world_size(group)
        # NCCL averages in the collective; gloo (CPU tests) sums and the scale is applied to each run afterwards
        self.avg = self.cuda and dist.get_backend(group) == "nccl"

    def _reduce(self, runs) -> None:
        flat = self.eng.grad_flat
        for off, ln in runs:
            view = flat[off:off + ln]
            if self.avg:
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(view, group=self.group)
                view.mul_(1.0 / self.world)

    def backward_and_reduce(self, d_out: torch.Tensor) -> None:
        self.eng.d_out.copy_(d_out)
        if not self.cuda:
            for a, b, runs in self.plan:
                self.run_range(a, b)
                self._reduce(runs)
            return
        main = torch.cuda.current_stream(self.eng.grad_flat.device)
        self.side.wait_stream(main)                      # the previous step's optimizer reads of grad_flat are ordered before us
        for a, b, runs in self.plan:
            self.run_range(a, b)
            if runs:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(ev)
                    self._reduce(runs)
        main.wait_stream(self.side)

"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (promptir_b200/ddp.py) on the emulated training program.
Two ranks each run half of a batch; after ONE all-reduce of the flat gradient buffer both hold the gradient of the
full-batch mean loss (what DDP computes for train.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emulator
from oracle import promptir_oracle as O
from promptir_b200 import PromptIR, ddp
from promptir_b200.train_engine import TrainEngine

B, H, W = 2, 32, 32


def _data():
    x, _ = O.synthetic_batch(B, H, W, seed=3)
    torch.manual_seed(11)
    return x, torch.randn(B, 3, H, W) / (3 * H * W)          # dL/d(out) of a per-sample-mean loss, one row per image


def _grads(model, x, d_out):
    eng = TrainEngine(model, x.shape[0], H, W, "cpu", torch.float32)
    emulator.run_train(eng, x, d_out)
    return eng


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    torch.manual_seed(0)
    model = PromptIR(decoder=True)
    x, d = _data()
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    eng = _grads(model, x[sl], d[sl] / per)                   # each rank: mean loss over ITS shard (as under DDP)
    ddp.attach_flat_grads(model, eng)
    ddp.allreduce_gradients(eng)
    # the overlapped exchange (backward in 4 pieces, finished gradient runs reduced after each piece) gives the same buffer
    one_shot = eng.grad_flat.clone()
    red = ddp.OverlappedReducer(eng, segments=4, run_range=lambda a, b: [emulator.DISPATCH[r["kind"]](r) for r in eng.bwd_ops[a:b]])
    covered = sum(ln for _, _, runs in red.plan for _, ln in runs)
    live = sum(eng.grads[n].numel() for n in eng.live_params)
    red.backward_and_reduce(d[sl] / per)
    overlap_err = ((eng.grad_flat - one_shot).norm() / one_shot.norm()).item()
    dead = [n for n, p in model.named_parameters() if p.grad is None]
    if rank == 0:
        ref = _grads(model, x, d / B).grad_flat               # single process, mean loss over the whole batch
        q.put((((eng.grad_flat - ref).norm() / ref.norm()).item(), len(dead),
               model.output.weight.grad.data_ptr() == eng.grads["output.weight"].data_ptr(), overlap_err, covered == live, len(red.plan)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_gloo_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rel, ndead, zero_copy, overlap_err, all_covered, pieces = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert rel < 1e-4 and ndead == 6 and zero_copy
    assert overlap_err < 1e-6 and all_covered and pieces == 4, (overlap_err, all_covered, pieces)

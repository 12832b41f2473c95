"""CPU: the batched / sharded tile scheduler (promptir_b200/tiling.py) -- single process, and world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import promptir_oracle as O
from promptir_b200 import tiling


def fake_model(t):                      # any per-tile function with no cross-sample coupling
    return t * 1.3 - 0.1 + 0.05 * t.flip(-2)


def cpu_blend(tiles, ys, xs, h, w):     # reference arithmetic of demo.py:43-47 (the CUDA kernel is tested under -m gpu)
    tile = tiles.shape[-1]
    acc = torch.zeros(tiles.shape[1], h, w)
    hit = torch.zeros_like(acc)
    i = 0
    for y in ys:
        for x in xs:
            acc[:, y:y + tile, x:x + tile] += tiles[i]
            hit[:, y:y + tile, x:x + tile] += 1
            i += 1
    return (acc / hit).clamp(0, 1)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 170, 171):
        for world in (1, 2, 3, 8):
            spans = [tiling.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    assert len(tiling.tile_origins(2160, 256, 32)) * len(tiling.tile_origins(3840, 256, 32)) == 170


def test_pad_input_matches_reference_rule():
    x = torch.rand(1, 3, 70, 52)
    a, h, w = tiling.pad_input(x, 8)
    b, h2, w2 = O.pad_to_multiple(x, 8)
    assert torch.equal(a, b) and (h, w) == (h2, w2) == (70, 52) and a.shape[-2:] == (72, 56)
    assert tiling.pad_input(torch.rand(1, 3, 64, 64))[0].shape[-2:] == (64, 64)


def test_batched_tile_eval_single_process():
    x = torch.rand(2, 3, 72, 56)
    ref = O.tiled_restore(fake_model, x, 32, 8)
    out = tiling.tile_eval(fake_model, x, 32, 8, batch=5, blend=cpu_blend)
    assert torch.allclose(out, ref, atol=1e-6)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    x = torch.rand(1, 3, 72, 56)
    out = tiling.tile_eval(fake_model, x, 32, 8, batch=4, blend=cpu_blend)
    if rank == 0:
        q.put(out.numpy())              # by value: a tensor travels as a shared-memory handle that dies with this process
    dist.barrier()
    dist.destroy_process_group()


def test_tile_eval_sharded_over_two_gloo_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = torch.from_numpy(q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    x = torch.rand(1, 3, 72, 56)
    assert torch.allclose(out, O.tiled_restore(fake_model, x, 32, 8), atol=1e-6)

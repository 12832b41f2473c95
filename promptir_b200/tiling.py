"""Batched, multi-GPU tiled inference: the B200 replacement for the sequential loop of demo.py:26-48.

`tile_eval(model, input_, tile, tile_overlap)` keeps the reference's signature and result (same tile origins,
same hit-count averaging, same clamp) but
  * gathers all tiles of the frame into batches and runs them through the engine `batch` tiles at a time,
  * with torch.distributed initialised, shards the tile list across ranks (contiguous, balanced; tiles are
    independent so there is no data-path collective) and all-gathers the restored tiles,
  * blends with one deterministic gather kernel (pir_tile_blend) instead of per-tile read-modify-write.
The reference's own demo.tile_eval also keeps working unchanged on the drop-in module (b=1 per call).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.nn.functional as F


def pad_input(input_: torch.Tensor, img_multiple_of: int = 8):
    """demo.py:17-24 (reflect-pad bottom/right to the next multiple; returns the padded tensor and the original size)."""
    height, width = input_.shape[2], input_.shape[3]
    padh = (height + img_multiple_of) // img_multiple_of * img_multiple_of - height if height % img_multiple_of else 0
    padw = (width + img_multiple_of) // img_multiple_of * img_multiple_of - width if width % img_multiple_of else 0
    return F.pad(input_, (0, padw, 0, padh), "reflect"), height, width


def tile_origins(extent: int, tile: int, overlap: int) -> List[int]:
    """demo.py:32-34."""
    return list(range(0, extent - tile, tile - overlap)) + [extent - tile]


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split of n items: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_tiles(frame: torch.Tensor, ys: List[int], xs: List[int], tile: int) -> torch.Tensor:
    """frame [C,H,W] -> [len(ys)*len(xs), C, tile, tile] in the reference's loop order (rows outer, columns inner)."""
    return torch.stack([frame[:, y:y + tile, x:x + tile] for y in ys for x in xs])


def _cuda_blend(tiles: torch.Tensor, ys: List[int], xs: List[int], h: int, w: int) -> torch.Tensor:
    from . import ops
    out = torch.empty(tiles.shape[1], h, w, dtype=torch.float32, device=tiles.device)
    ops.tile_blend(tiles.contiguous(), torch.tensor(ys, dtype=torch.int32, device=tiles.device),
                   torch.tensor(xs, dtype=torch.int32, device=tiles.device), out, torch.cuda.current_stream().cuda_stream)
    return out


def tile_eval(model: Callable[[torch.Tensor], torch.Tensor], input_: torch.Tensor, tile: int = 128, tile_overlap: int = 32,
              batch: int = 16, group=None, blend: Optional[Callable] = None) -> torch.Tensor:
    """Drop-in for demo.tile_eval with batching and optional sharding over a torch.distributed process group."""
    import torch.distributed as dist
    b, c, h, w = input_.shape
    tile = min(tile, h, w)
    assert tile % 8 == 0, "tile size should be multiple of 8"
    ys, xs = tile_origins(h, tile, tile_overlap), tile_origins(w, tile, tile_overlap)
    blend = blend or _cuda_blend
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    outs = []
    for i in range(b):
        tiles = gather_tiles(input_[i], ys, xs, tile)
        n = tiles.shape[0]
        lo, hi = shard_bounds(n, rank, world)
        mine = [model(tiles[s:min(s + batch, hi)]) for s in range(lo, hi, batch)]
        mine = torch.cat(mine) if mine else tiles.new_zeros((0, c, tile, tile))
        if world > 1:
            cap = shard_bounds(n, 0, world)[1]                      # largest shard
            padded = torch.zeros((cap,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
            padded[:mine.shape[0]] = mine
            parts = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(parts, padded, group=group)
            mine = torch.cat([p[:shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0]] for r, p in enumerate(parts)])
        outs.append(blend(mine.float(), ys, xs, h, w))
    return torch.stack(outs).to(input_.dtype)

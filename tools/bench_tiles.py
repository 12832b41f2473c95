"""cfg3 (BASELINE.json configs[2]): tiled inference of a synthetic 3840x2160 frame, tile 256, overlap 32 (170 tiles),
batched through the engine; with torchrun the tile list is sharded over the ranks.  Diagnostic / scaling check.

    python tools/bench_tiles.py [frames] [batch]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_tiles.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptir_b200 import PromptIR, tiling  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = PromptIR(decoder=True).eval().to(dev)
    g = torch.Generator().manual_seed(1)
    frame = torch.rand(1, 3, 2160, 3840, generator=g).to(dev)
    x, h, w = tiling.pad_input(frame, 8)
    with torch.no_grad():
        for _ in range(2):
            out = tiling.tile_eval(model, x, 256, 32, batch=batch)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            out = tiling.tile_eval(model, x, 256, 32, batch=batch)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        print(json.dumps({"workload": "3840x2160 frame, tile 256, overlap 32, 170 tiles", "n_gpus": world, "batch": batch, "ms_per_frame": ms,
                          "frame_MP_per_s": 3840 * 2160 / 1e6 / ms * 1e3, "tile_MP_per_s": 170 * 256 * 256 / 1e6 / ms * 1e3,
                          "finite": bool(torch.isfinite(out).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

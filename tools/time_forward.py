"""Per-op CUDA-event timing of one engine program (eager launches), grouped by kernel tag, plus whole-forward
eager and CUDA-graph times.  Diagnostic tool for gpurun; not the benchmark (see bench.py).

    python tools/time_forward.py [B H W dtype]
"""
import json
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptir_b200 import PromptIR  # noqa: E402
from promptir_b200.engine import Engine  # noqa: E402


def main():
    B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 256, 256)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[4] if len(sys.argv) > 4 else "bf16"]
    torch.manual_seed(0)
    m = PromptIR(decoder=True).eval().cuda()
    eng = Engine(m, B, H, W, "cuda", dt)
    eng.img_in.copy_(torch.rand_like(eng.img_in))
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        eng.launch_all(s)
    torch.cuda.synchronize()
    # per-op
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in eng.ops]
    for (a, b), r in zip(evs, eng.ops):
        a.record()
        r["launch"](s)
        b.record()
    torch.cuda.synchronize()
    agg, cnt, byt = defaultdict(float), defaultdict(int), defaultdict(float)
    rows = []
    for (a, b), r in zip(evs, eng.ops):
        ms = a.elapsed_time(b)
        key = r.get("tag") or r["kind"]
        agg[key] += ms
        cnt[key] += 1
        nbytes = 0
        for k in ("a", "x", "qkv"):
            if k in r and torch.is_tensor(r[k]) and r[k].dim() == 4:
                nbytes += r[k].shape[0] * r[k].shape[1] * r[k].shape[2] * r[k].shape[3] * r[k].element_size()
        if "out" in r and torch.is_tensor(r["out"]):
            nbytes += r["out"].numel() * r["out"].element_size()
        if r.get("res") is not None:
            nbytes += r["res"].numel() * r["res"].element_size()
        if r["kind"] == "mdta_gram":
            nbytes = nbytes * 2 // 3
        byt[key] += nbytes
        shape = tuple(r["out"].shape) if "out" in r and torch.is_tensor(r["out"]) else ()
        rows.append((ms, key, shape, nbytes))
    tot = sum(agg.values())
    print(f"B={B} H={H} W={W} {dt}: sum of per-op times {tot:.3f} ms over {len(eng.ops)} ops")
    for k in sorted(agg, key=lambda k: -agg[k]):
        gbs = byt[k] / agg[k] / 1e6 if agg[k] > 0 else 0
        print(f"  {k:14s} n={cnt[k]:4d}  {agg[k]:8.3f} ms  {100 * agg[k] / tot:5.1f}%   ~{gbs:7.0f} GB/s (activation bytes only)")
    print("slowest ops:")
    for ms, key, shape, nb in sorted(rows, key=lambda r: -r[0])[:25]:
        print(f"  {ms:7.3f} ms {key:12s} {shape}  ~{nb / ms / 1e6 if ms else 0:6.0f} GB/s")
    # whole forward
    def timed(fn, n=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    te = timed(lambda: eng.replay(False))
    tg = timed(lambda: eng.replay(True))
    from promptir_b200.engine import SplitEngine
    sp = SplitEngine(m, B, H, W, "cuda", dt)
    sp.img_in.copy_(eng.img_in)
    ts = timed(lambda: sp.replay(True))
    print(f"split graph {ts:.3f} ms ({B * H * W / 1e6 / ts * 1e3:.1f} MP/s)   max |split - whole| = {(sp.out - eng.out).abs().max().item():.3e}")
    mp = B * H * W / 1e6
    print(f"eager {te:.3f} ms ({mp / te * 1e3:.1f} MP/s)   graph {tg:.3f} ms ({mp / tg * 1e3:.1f} MP/s)")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/time_forward.json", "w") as f:
        json.dump({"B": B, "H": H, "W": W, "eager_ms": te, "graph_ms": tg, "by_tag_ms": dict(agg),
                   "ops": [{"i": i, "ms": r[0], "tag": r[1], "shape": list(r[2]), "bytes": r[3]} for i, r in enumerate(rows)]}, f)


if __name__ == "__main__":
    main()

"""Run ONE full-size TransformerBlock's kernels inside a cudaProfiler range (for `ncu --profile-from-start off`).

    python tools/profile_block.py [B H W dtype stage_tag_index]
Default: B=16, 256x256, bf16, the first refinement block (C=96 at full resolution: the shapes that dominate the step).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptir_b200 import PromptIR  # noqa: E402
from promptir_b200.engine import Engine  # noqa: E402


def main():
    B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 256, 256)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[4] if len(sys.argv) > 4 else "bf16"]
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 43            # K1-launch ordinal: 43 = refinement[0]
    torch.manual_seed(0)
    m = PromptIR(decoder=True).eval().cuda()
    eng = Engine(m, B, H, W, "cuda", dt)
    eng.img_in.copy_(torch.rand_like(eng.img_in))
    s = torch.cuda.current_stream().cuda_stream
    eng.launch_all(s)
    torch.cuda.synchronize()
    k1 = [i for i, r in enumerate(eng.ops) if r.get("tag") in ("K1", "K12")]
    lo = k1[which]
    hi = k1[which + 1] if which + 1 < len(k1) else lo + 9
    block = [r for r in eng.ops[lo:hi] if r.get("tag", "").startswith("K")]
    for r in block:                      # warm
        r["launch"](s)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for r in block:
        r["launch"](s)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", [r["tag"] for r in block], "C =", eng.ops[lo]["a"].shape[-1])


if __name__ == "__main__":
    main()

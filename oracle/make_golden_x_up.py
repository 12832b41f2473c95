"""Generate tests/golden/xrestormer_seed0_up.npz by running the REAL reference net/prompt_xrestormer.py (dev container only; imported
from /root/reference by path, never copied; `torchstat`, an import-only dependency, is stubbed).

The cases of make_golden_x.py (64, 64x128, 128) never resize a prompt UP: PromptBlock's prompts are 64 / 32 / 16 wide at H/2, H/4,
H/8 (prompt_xrestormer.py:421-431), so F.interpolate(..., align_corners=True) is a down-scale or the identity there.  BASELINE.json
configs[4] runs at 512x512 where they are up-sampled; these two cases pin that branch (x2 at 256x256, x1.5 / x2.5 at 192x320).
Only the outputs are stored (the inputs are oracle.promptir_oracle.synthetic_batch(b, h, w, seed), whose digest is stored beside).

    python oracle/make_golden_x_up.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle.make_golden_x import import_reference  # noqa: E402
from oracle.promptir_oracle import synthetic_batch  # noqa: E402

CASES = {"x256": (1, 256, 256, 4), "x192x320": (1, 192, 320, 6)}


def main():
    PromptXRestormer = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    ref = PromptXRestormer().eval()
    blob = {}
    for name, (b, h, w, seed) in CASES.items():
        x, _ = synthetic_batch(b, h, w, seed=seed)
        with torch.no_grad():
            y = ref(x)
        blob[name + "_out"] = y.numpy()
        d = x.double()
        blob[name + "_in_digest"] = np.array([d.sum().item(), d.abs().sum().item(), d.flatten()[0].item(), d.flatten()[-1].item()])
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    np.savez_compressed(os.path.join(OUT, "xrestormer_seed0_up.npz"), **blob)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

"""Generate tests/golden/forward_seed0_up.npz + taps_seed0_up.json by running the REAL reference (dev container only; the
reference is imported from /root/reference, never copied).

The cases of make_golden.py are at most 128x128, where every PromptGenBlock either resizes its prompt DOWN (32, 40x24, 64) or not
at all (128: S == H/2, model.py:231 is the identity).  BASELINE.json's headline configuration runs at 256x256, where all three
prompts are bilinearly UP-sampled (x2); these cases pin that branch -- the exact x2 and non-integer factors (192x320: x1.5 / x2.5)
of F.interpolate(..., mode="bilinear") with align_corners=False -- in the oracle and in the CUDA path.

    python oracle/make_golden_256.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle.make_golden import hooked_forward, tap_digest  # noqa: E402
from oracle.promptir_oracle import synthetic_batch  # noqa: E402

CASES = {"cfg2_256": (1, 256, 256, 4), "a192x320": (1, 192, 320, 6)}


def import_reference():
    """The reference's net/model.py by file path (the name `net.model` may already be taken by this repo's import shim)."""
    import importlib.util
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_net_model", "/root/reference/net/model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PromptIR


def main():
    PromptIR = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    ref = PromptIR(decoder=True).eval()
    blob, taps_meta = {}, {}
    for name, (b, h, w, seed) in CASES.items():
        x, _ = synthetic_batch(b, h, w, seed=seed)
        y, taps = hooked_forward(ref, x)
        blob[name + "_in"], blob[name + "_out"] = x.numpy(), y.numpy()
        taps_meta[name] = {k: tap_digest(v) for k, v in taps.items()}
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    np.savez_compressed(os.path.join(OUT, "forward_seed0_up.npz"), **blob)
    with open(os.path.join(OUT, "taps_seed0_up.json"), "w") as f:
        json.dump(taps_meta, f)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

"""Generate tests/golden/xrestormer_seed0.npz by running the REAL reference net/prompt_xrestormer.py (dev container only; the
reference is imported from /root/reference by path, never copied; `torchstat` -- an import-only dependency -- is stubbed).

    python oracle/make_golden_x.py
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle.make_golden import param_digest  # noqa: E402
from oracle.promptir_oracle import synthetic_batch  # noqa: E402


def import_reference():
    sys.dont_write_bytecode = True
    sys.modules.setdefault("torchstat", types.SimpleNamespace(stat=None))
    spec = importlib.util.spec_from_file_location("ref_prompt_xrestormer", "/root/reference/net/prompt_xrestormer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PromptXRestormer


def main():
    PromptXRestormer = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    ref = PromptXRestormer().eval()
    sd = ref.state_dict()
    digest = param_digest(sd)
    some = [k for k in sd if any(t in k for t in ("rel_pos_emb", "prompt_param", "temperature", "output", "patch_embed", "prompt3.conv."))]
    meta = {"torch": torch.__version__, "seed": 0, "keys": list(sd.keys()), "n_params": sum(p.numel() for p in ref.parameters()),
            "params": {k: digest[k] for k in some}}
    with open(os.path.join(OUT, "xrestormer_params_seed0.json"), "w") as f:
        json.dump(meta, f)
    blob = {}
    for name, (b, h, w, seed) in {"x64": (1, 64, 64, 1), "x64x128": (2, 64, 128, 2), "x128": (1, 128, 128, 3)}.items():
        x, _ = synthetic_batch(b, h, w, seed=seed)
        with torch.no_grad():
            y = ref(x)
        blob[name + "_in"], blob[name + "_out"] = x.numpy(), y.numpy()
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    np.savez_compressed(os.path.join(OUT, "xrestormer_seed0.npz"), **blob)


if __name__ == "__main__":
    main()

"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Dev-container only: needs /root/reference (read-only).  The reference is imported, never copied.  The
fixtures pin (1) seed-identical parameter construction, (2) the reference forward on small synthetic
inputs, including named intermediates, (3) demo.py's pad_input / tile_eval.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz, *.json
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle.promptir_oracle import synthetic_batch  # noqa: E402


def import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    # demo.py needs lightning + matplotlib at import time; stub them
    lit = types.ModuleType("lightning")
    litp = types.ModuleType("lightning.pytorch")
    litp.LightningModule = torch.nn.Module
    lit.pytorch = litp
    sys.modules.setdefault("lightning", lit)
    sys.modules.setdefault("lightning.pytorch", litp)
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    from net.model import PromptIR  # the reference's
    import demo
    return PromptIR, demo


def param_digest(sd):
    """Cheap, order-sensitive fingerprint of every tensor: (shape, sum, abs-sum, first, last) in float64."""
    out = {}
    for k, v in sd.items():
        d = v.detach().double().flatten()
        out[k] = {"shape": list(v.shape), "sum": d.sum().item(), "abs": d.abs().sum().item(),
                  "first": d[0].item(), "last": d[-1].item()}
    return out


def tap_digest(t):
    d = t.detach().double()
    flat = d.flatten()
    idx = torch.linspace(0, flat.numel() - 1, 64).long()
    return {"shape": list(t.shape), "mean": d.mean().item(), "abs": d.abs().mean().item(),
            "sample": flat[idx].tolist()}


def hooked_forward(model, x):
    names = ["patch_embed", "encoder_level1", "encoder_level2", "encoder_level3", "latent", "prompt3",
             "reduce_noise_level3", "decoder_level3", "prompt2", "reduce_noise_level2", "decoder_level2",
             "prompt1", "reduce_noise_level1", "decoder_level1", "refinement"]
    taps, hs = {}, []
    for n in names:
        hs.append(getattr(model, n).register_forward_hook(lambda m, i, o, n=n: taps.__setitem__(n, o)))
    with torch.no_grad():
        y = model(x)
    for h in hs:
        h.remove()
    return y, taps


def main():
    os.makedirs(OUT, exist_ok=True)
    PromptIR, demo = import_reference()
    torch.set_num_threads(os.cpu_count())

    # ---- case A: default architecture, seed 0 -------------------------------------------------------
    torch.manual_seed(0)
    ref = PromptIR(decoder=True).eval()
    meta = {"torch": torch.__version__, "seed": 0, "params": param_digest(ref.state_dict()),
            "n_params": sum(p.numel() for p in ref.parameters()), "keys": list(ref.state_dict().keys())}
    with open(os.path.join(OUT, "params_seed0.json"), "w") as f:
        json.dump(meta, f)

    cases = {"a32": (1, 32, 32, 1), "a40x24": (2, 40, 24, 2), "a64": (1, 64, 64, 3), "cfg1_128": (1, 128, 128, 1)}
    blob, taps_meta = {}, {}
    for name, (b, h, w, seed) in cases.items():
        x, _ = synthetic_batch(b, h, w, seed=seed)
        y, taps = hooked_forward(ref, x)
        blob[name + "_in"] = x.numpy()
        blob[name + "_out"] = y.numpy()
        taps_meta[name] = {k: tap_digest(v) for k, v in taps.items()}
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    # ---- demo.py helpers ---------------------------------------------------------------------------
    x, _ = synthetic_batch(1, 70, 52, seed=7)
    xp, h0, w0 = demo.pad_input(x, 8)
    with torch.no_grad():
        tiled = demo.tile_eval(ref, xp, tile=32, tile_overlap=8)
    blob["tile_in"] = x.numpy()
    blob["tile_padded"] = xp.numpy()
    blob["tile_out"] = tiled.numpy()
    np.savez_compressed(os.path.join(OUT, "forward_seed0.npz"), **blob)
    with open(os.path.join(OUT, "taps_seed0.json"), "w") as f:
        json.dump(taps_meta, f)

    # ---- case B: bias=True + BiasFree LayerNorm, seed 3 ----------------------------------------------
    torch.manual_seed(3)
    refb = PromptIR(decoder=True, bias=True, LayerNorm_type="BiasFree").eval()
    x, _ = synthetic_batch(1, 32, 48, seed=5)
    with torch.no_grad():
        y = refb(x)
    np.savez_compressed(os.path.join(OUT, "forward_seed3_biasfree.npz"), x=x.numpy(), y=y.numpy())
    with open(os.path.join(OUT, "params_seed3_biasfree.json"), "w") as f:
        json.dump({"seed": 3, "params": param_digest(refb.state_dict())}, f)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

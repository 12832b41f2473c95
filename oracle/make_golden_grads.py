"""Generate tests/golden/grads_seed0.json by running the REAL reference through the step train.py:37-46 makes
(restored = net(degrad_patch); loss = L1Loss()(restored, clean_patch); loss.backward()) with autograd -- dev container only; the
reference's net/model.py and net/prompt_xrestormer.py are imported from /root/reference by path, never copied.

The training-path tests compare the hand-derived backward programs with autograd of the ORACLE; this fixture pins that oracle
autograd to the real module's: loss value and, per parameter, a digest of the gradient (shape, sum, abs-sum, L2 norm, 16 samples).
Parameters are the seed-0 construction with temperature / LayerNorm affine perturbed (as in the tests: at their init values 1 / 1 / 0
mistakes in those gradients hide), perturbation drawn from a fixed generator so the test can rebuild it.

    python oracle/make_golden_grads.py
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle.make_golden_256 import import_reference  # noqa: E402
from oracle.make_golden_x import import_reference as import_reference_x  # noqa: E402
from oracle.promptir_oracle import synthetic_batch  # noqa: E402


def perturb(named_parameters, seed: int = 11) -> None:
    """In place, in registration order, from one CPU generator (the test applies the same rule to the drop-in module)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in named_parameters:
            if n.endswith("temperature") or ("norm" in n and n.endswith("weight")):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif "norm" in n and n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)


def grad_digest(g):
    if g is None:
        return None
    d = g.detach().double().flatten()
    idx = torch.linspace(0, d.numel() - 1, min(16, d.numel())).long()
    return {"shape": list(g.shape), "sum": d.sum().item(), "abs": d.abs().sum().item(), "norm": d.norm().item(), "sample": d[idx].tolist()}


def run(module, shape, seed):
    x, clean = synthetic_batch(*shape, seed=seed)
    module.zero_grad(set_to_none=True)
    loss = torch.nn.L1Loss()(module(x), clean)                  # train.py:28,41-43
    loss.backward()
    return {"shape": list(shape), "seed": seed, "loss": loss.item(),
            "grads": {n: grad_digest(p.grad) for n, p in module.named_parameters()}}


def main():
    torch.set_num_threads(os.cpu_count())
    out = {"torch": torch.__version__, "perturb_seed": 11}
    PromptIR = import_reference()
    torch.manual_seed(0)
    ref = PromptIR(decoder=True).train()
    perturb(ref.named_parameters())
    out["promptir"] = {"a32": run(ref, (2, 32, 32), 7), "a40x24": run(ref, (1, 40, 24), 8)}
    PromptXRestormer = import_reference_x()
    torch.manual_seed(0)
    refx = PromptXRestormer(num_blocks=[1, 1, 1, 2], num_refinement_blocks=1).train()
    perturb(refx.named_parameters())
    out["xrestormer_small"] = {"x64": run(refx, (1, 64, 64), 7)}
    with open(os.path.join(OUT, "grads_seed0.json"), "w") as f:
        json.dump(out, f)
    for k in ("promptir", "xrestormer_small"):
        for c, r in out[k].items():
            print(k, c, "loss", r["loss"], "params with grad", sum(v is not None for v in r["grads"].values()), "/", len(r["grads"]))


if __name__ == "__main__":
    main()

"""CPU: autograd of the oracles against gradients of the REAL reference modules (tests/golden/grads_seed0.json, written by
oracle/make_golden_grads.py, which runs the step train.py:37-46 makes through the unmodified net/model.py and
net/prompt_xrestormer.py).  The training-path tests (CPU emulator and B200) compare the hand-derived backward programs with autograd of
the oracle; this is what pins that oracle autograd -- loss and every parameter gradient -- to the reference's."""
import json
import math
import os

import pytest
import torch

from oracle import promptir_oracle as O
from oracle import xrestormer_oracle as XO
from oracle.make_golden_grads import perturb
from promptir_b200 import PromptIR, PromptXRestormer


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "grads_seed0.json")))


def _check(rec, sd, forward):
    x, clean = O.synthetic_batch(*rec["shape"], seed=rec["seed"])
    loss = torch.nn.functional.l1_loss(forward(sd, x), clean)
    loss.backward()
    assert abs(loss.item() - rec["loss"]) <= 1e-6 * abs(rec["loss"])
    assert list(sd.keys()) == list(rec["grads"].keys())
    worst = 0.0
    for n, d in rec["grads"].items():
        g = sd[n].grad
        if d is None:                                             # the six convs forward() never uses: no gradient
            assert g is None, n
            continue
        assert g is not None and list(g.shape) == d["shape"], n
        f = g.double().flatten()
        scale = max(d["norm"], 1e-30)
        rms = scale / math.sqrt(f.numel())
        idx = torch.linspace(0, f.numel() - 1, min(16, f.numel())).long()
        errs = (abs(f.norm().item() - d["norm"]) / scale, abs(f.abs().sum().item() - d["abs"]) / max(d["abs"], 1e-30),
                abs(f.sum().item() - d["sum"]) / max(d["abs"], 1e-30),
                (f[idx] - torch.tensor(d["sample"], dtype=torch.float64)).abs().max().item() / max(rms, 1e-30) * 1e-2)
        worst = max(worst, *errs)
        assert max(errs) <= 2e-4, (n, errs)
    return worst


@pytest.mark.parametrize("case", ["a32", "a40x24"])
def test_promptir_oracle_autograd_matches_reference_gradients(golden, case):
    torch.manual_seed(0)
    m = PromptIR(decoder=True)
    perturb(m.named_parameters(), golden["perturb_seed"])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    worst = _check(golden["promptir"][case], sd, O.promptir_forward)
    assert sum(v is None for v in golden["promptir"][case]["grads"].values()) == 6
    print(f"[oracle grads] promptir {case}: worst relative digest error {worst:.2e}")


def test_xrestormer_oracle_autograd_matches_reference_gradients(golden):
    torch.manual_seed(0)
    m = PromptXRestormer(num_blocks=[1, 1, 1, 2], num_refinement_blocks=1)
    perturb(m.named_parameters(), golden["perturb_seed"])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    worst = _check(golden["xrestormer_small"]["x64"], sd,
                   lambda s, x: XO.xrestormer_forward(s, x, num_blocks=(1, 1, 1, 2), num_refinement_blocks=1))
    print(f"[oracle grads] xrestormer (1,1,1,2 blocks) x64: worst relative digest error {worst:.2e}")

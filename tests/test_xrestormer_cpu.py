"""CPU: PromptXRestormer (net/prompt_xrestormer.py) -- oracle pinned to the real reference's golden outputs,
seed-identical module construction, and the XEngine program interpreted by the torch emulator."""
import json
import os

import numpy as np
import pytest
import torch

import emulator
from oracle import xrestormer_oracle as XO
from promptir_b200 import PromptXRestormer
from promptir_b200.xengine import XEngine


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return PromptXRestormer().eval()


def test_seed_identical_parameters(model, golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "xrestormer_params_seed0.json")))
    sd = model.state_dict()
    assert list(sd.keys()) == meta["keys"] and sum(p.numel() for p in model.parameters()) == meta["n_params"]
    for k, d in meta["params"].items():
        v = sd[k].double().flatten()
        assert list(sd[k].shape) == d["shape"]
        assert abs(v.sum().item() - d["sum"]) <= 1e-9 * max(1.0, abs(d["sum"])) and v[0].item() == d["first"] and v[-1].item() == d["last"], k


def test_oracle_matches_reference_golden(model, golden_dir):
    g = np.load(os.path.join(golden_dir, "xrestormer_seed0.npz"))
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for case in ("x64", "x64x128"):
            y = XO.xrestormer_forward(sd, torch.from_numpy(g[case + "_in"]))
            assert (y - torch.from_numpy(g[case + "_out"])).abs().max().item() <= 5e-6


@pytest.mark.parametrize("dt,lim", [(torch.float32, 3e-5), (torch.bfloat16, 3e-2)])
def test_program_emulation_matches_reference(model, golden_dir, dt, lim):
    g = np.load(os.path.join(golden_dir, "xrestormer_seed0.npz"))
    x, yref = torch.from_numpy(g["x64x128_in"]), torch.from_numpy(g["x64x128_out"])
    eng = XEngine(model, x.shape[0], x.shape[2], x.shape[3], "cpu", dt)
    y = emulator.run_program(eng, x)
    assert (y.clamp(0, 1) - yref.clamp(0, 1)).abs().max().item() <= lim
    kinds = [r["kind"] for r in eng.ops]
    assert kinds.count("ocab") == 4 + 6 + 6 + 8 + 6 + 6 + 4 + 4 + 3 and kinds.count("mdta_gram") == kinds.count("ocab")


def test_errors_are_loud(model):
    with pytest.raises(RuntimeError):
        model(torch.rand(1, 3, 64, 64))                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        XEngine(model, 1, 96, 64, "cpu", torch.float32)       # not a multiple of 64

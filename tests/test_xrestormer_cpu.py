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


@pytest.mark.parametrize("case,shape,seed", [("x256", (1, 256, 256), 4), ("x192x320", (1, 192, 320), 6)])
def test_oracle_matches_reference_golden_with_upsampled_prompts(model, golden_dir, case, shape, seed):
    """The goldens above never resize a prompt UP (prompts are 64 / 32 / 16 wide at H/2, H/4, H/8); BASELINE.json configs[4] runs at
    512x512 where PromptBlock's F.interpolate(align_corners=True) up-samples.  oracle/make_golden_x_up.py ran the real reference at
    256x256 (x2) and 192x320 (x1.5 / x2.5); only its outputs are stored, the inputs are regenerated and checked by digest."""
    from oracle.promptir_oracle import synthetic_batch
    g = np.load(os.path.join(golden_dir, "xrestormer_seed0_up.npz"))
    x, _ = synthetic_batch(*shape, seed=seed)
    d = x.double()
    digest = np.array([d.sum().item(), d.abs().sum().item(), d.flatten()[0].item(), d.flatten()[-1].item()])
    assert np.array_equal(digest, g[case + "_in_digest"]), "synthetic_batch no longer reproduces the golden input"
    with torch.no_grad():
        y = XO.xrestormer_forward({k: v.detach() for k, v in model.state_dict().items()}, x)
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


def test_backward_program_matches_autograd():
    """XTrainEngine (forward + hand-derived backward, incl. the OCAB backward and the PromptBlock wiring) emulated in fp32 vs autograd."""
    from promptir_b200.xtrain_engine import XTrainEngine
    from oracle import promptir_oracle as O
    torch.manual_seed(0)
    m = PromptXRestormer(num_blocks=[1, 1, 1, 2], num_refinement_blocks=1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("temperature"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("bias"):
                p.copy_(torch.randn_like(p) * 0.2)
    B, H, W = 1, 64, 128
    x, _ = O.synthetic_batch(B, H, W, seed=3)
    torch.manual_seed(5)
    d_out = torch.randn(B, 3, H, W) / (3 * H * W)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xin = x.clone().requires_grad_(True)
    ref_out = XO.xrestormer_forward(sd, xin, num_blocks=(1, 1, 1, 2), num_refinement_blocks=1)
    ref_out.backward(d_out)
    eng = XTrainEngine(m, B, H, W, "cpu", torch.float32, input_grad=True)
    out, grads = emulator.run_train(eng, x, d_out)
    assert (out - ref_out.detach()).abs().max().item() < 3e-5
    assert all(v.grad is not None for v in sd.values())               # no dead parameters in this network
    worst = max((((grads[n] - v.grad).norm() / v.grad.norm().clamp_min(1e-30)).item(), n) for n, v in sd.items())
    assert worst[0] < 3e-4, worst
    assert ((eng.d_img - xin.grad).norm() / xin.grad.norm()).item() < 3e-4

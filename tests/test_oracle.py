"""CPU: the oracle against the golden vectors produced by the REAL reference (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import promptir_oracle as O
from promptir_b200 import PromptIR


@pytest.fixture(scope="module")
def seeded():
    torch.manual_seed(0)
    m = PromptIR(decoder=True).eval()
    return m, {k: v.detach() for k, v in m.state_dict().items()}


def test_seed_identical_parameters(seeded, golden_dir):
    """Same constructor order as net/model.py:260-320 -> same RNG stream -> identical tensors (digests from the reference)."""
    meta = json.load(open(os.path.join(golden_dir, "params_seed0.json")))
    _, sd = seeded
    assert list(sd.keys()) == meta["keys"] and len(sd) == 548
    assert sum(v.numel() for v in sd.values()) == meta["n_params"] == 35592263
    for k, v in sd.items():
        d = meta["params"][k]
        assert list(v.shape) == d["shape"], k
        f = v.double().flatten()
        assert abs(f.sum().item() - d["sum"]) <= 1e-9 * max(1, abs(d["sum"])), k
        assert f[0].item() == d["first"] and f[-1].item() == d["last"], k


@pytest.mark.parametrize("case", ["a32", "a40x24", "a64", "cfg1_128"])
def test_oracle_matches_reference_outputs(seeded, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "forward_seed0.npz"))
    taps = {}
    with torch.no_grad():
        y = O.promptir_forward(seeded[1], torch.from_numpy(g[case + "_in"]), taps=taps)
    assert (y - torch.from_numpy(g[case + "_out"])).abs().max().item() < 5e-6
    meta = json.load(open(os.path.join(golden_dir, "taps_seed0.json")))[case]
    for name, d in meta.items():
        t = taps[name].double()
        assert list(t.shape) == d["shape"], name
        flat = t.flatten()
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        assert (flat[idx] - torch.tensor(d["sample"], dtype=torch.float64)).abs().max().item() < 2e-5, name


@pytest.mark.parametrize("case", ["cfg2_256", "a192x320"])
def test_oracle_matches_reference_outputs_with_upsampled_prompts(seeded, golden_dir, case):
    """The cases above never resize a prompt UP (model.py:231: down at 32 / 40x24 / 64, identity at 128).  At the headline size
    256x256 all three prompts are up-sampled x2, at 192x320 by x1.5 / x2.5 (oracle/make_golden_256.py, real reference)."""
    g = np.load(os.path.join(golden_dir, "forward_seed0_up.npz"))
    taps = {}
    with torch.no_grad():
        y = O.promptir_forward(seeded[1], torch.from_numpy(g[case + "_in"]), taps=taps)
    assert (y - torch.from_numpy(g[case + "_out"])).abs().max().item() < 5e-6
    meta = json.load(open(os.path.join(golden_dir, "taps_seed0_up.json")))[case]
    for name, d in meta.items():
        t = taps[name].double()
        assert list(t.shape) == d["shape"], name
        flat = t.flatten()
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        assert (flat[idx] - torch.tensor(d["sample"], dtype=torch.float64)).abs().max().item() < 2e-5, name


def test_oracle_biasfree_with_bias(golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_seed3_biasfree.npz"))
    torch.manual_seed(3)
    m = PromptIR(decoder=True, bias=True, LayerNorm_type="BiasFree").eval()
    arch = O.ArchSpec(layernorm_type="BiasFree")
    with torch.no_grad():
        y = O.promptir_forward({k: v.detach() for k, v in m.state_dict().items()}, torch.from_numpy(g["x"]), arch)
    assert (y - torch.from_numpy(g["y"])).abs().max().item() < 5e-6


def test_pad_and_tile_helpers_match_demo(seeded, golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_seed0.npz"))
    x = torch.from_numpy(g["tile_in"])
    xp, h, w = O.pad_to_multiple(x, 8)
    assert (h, w) == (70, 52) and torch.equal(xp, torch.from_numpy(g["tile_padded"]))
    with torch.no_grad():
        out = O.tiled_restore(lambda t: O.promptir_forward(seeded[1], t), xp, tile=32, overlap=8)
    assert (out - torch.from_numpy(g["tile_out"])).abs().max().item() < 5e-6
    assert O.tile_origins(2160, 256, 32) == list(range(0, 1904, 224)) + [1904] and len(O.tile_origins(3840, 256, 32)) == 17


def test_synthetic_batch_is_deterministic():
    a, ca = O.synthetic_batch(5, 32, 32, seed=4)
    b, cb = O.synthetic_batch(5, 32, 32, seed=4)
    assert torch.equal(a, b) and torch.equal(ca, cb) and 0 <= a.min() and a.max() <= 1


def test_gelu_sigmoid_polynomial_bound():
    """The CUDA kernels evaluate erf-GELU as x*sigmoid(2x(a + b u + c u^2)), u = min(x^2, 25) (csrc/common.cuh gelu_erf).
    Restated here in fp32 with the same constants: it must stay within 2.6e-5 of torch's exact erf GELU everywhere."""
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "promptir_b200", "csrc", "common.cuh")).read()
    a, b, c = (float(v) for v in re.search(r"kGeluA = ([-0-9.e]+)f, kGeluB = ([-0-9.e]+)f, kGeluC = ([-0-9.e]+)f", src).groups())
    x = torch.cat([torch.linspace(-30, 30, 2_000_001), torch.tensor([-1e4, -100.0, 100.0, 1e4, 0.0])])
    u = (x * x).clamp_max(25.0)
    t = u * (u * c + b) + a
    y = x / (1.0 + torch.exp2(x * t))
    ref = torch.nn.functional.gelu(x.double()).float()
    assert (y - ref).abs().max().item() <= 2.6e-5


def test_product_synthetic_data_is_the_oracles():
    """bench.py / tools draw their inputs from promptir_b200.synth; the oracle keeps its own copy (golden generation) -- keep them equal."""
    from promptir_b200 import synth
    a, b = O.synthetic_batch(5, 32, 40, seed=3), synth.synthetic_batch(5, 32, 40, seed=3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert O.psnr(a[0], a[1]) == synth.psnr(b[0], b[1])


def test_gate_backward_phi_approximation_bound():
    """pir_dwconv3x3(gate=2) evaluates Phi(x) with the forward gate's logistic-polynomial fit: |dPhi| <= 5.1e-5 everywhere."""
    x = torch.linspace(-12, 12, 200001, dtype=torch.float64)
    a, b, c = -2.301121339544986, -0.10677572399054727, 0.0010142630610895519          # kGeluA/B/C of csrc/common.cuh
    u = torch.clamp(x * x, max=25.0)
    approx = 1 / (1 + torch.exp2(x * (a + b * u + c * u * u)))
    exact = 0.5 * (1 + torch.erf(x * 0.7071067811865476))
    assert (approx - exact).abs().max().item() <= 5.1e-5

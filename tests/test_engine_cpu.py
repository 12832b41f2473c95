"""CPU: host logic of the product -- module/state_dict compatibility, weight packing identities, and the engine's
program (buffer plan, concat folding, op order) interpreted by the torch emulator against the reference goldens."""
import os

import numpy as np
import pytest
import torch

import emulator
from oracle import promptir_oracle as O
from promptir_b200 import PromptIR, packing
from promptir_b200.engine import Engine, op_cost


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return PromptIR(decoder=True).eval()


@pytest.mark.parametrize("dt,lim", [(torch.float32, 2e-5), (torch.float16, 2e-3), (torch.bfloat16, 2e-2)])
def test_program_emulation_matches_reference(model, golden_dir, dt, lim):
    g = np.load(os.path.join(golden_dir, "forward_seed0.npz"))
    for case in ("a32", "a40x24"):
        x, yref = torch.from_numpy(g[case + "_in"]), torch.from_numpy(g[case + "_out"])
        eng = Engine(model, x.shape[0], x.shape[2], x.shape[3], "cpu", dt)
        y = emulator.run_program(eng, x)
        assert (y.clamp(0, 1) - yref.clamp(0, 1)).abs().max().item() <= lim


def test_program_shape(model):
    eng = Engine(model, 2, 64, 64, "cpu", torch.bfloat16)
    kinds = [r["kind"] for r in eng.ops]
    fused = kinds.count("pwdw")                          # blocks with C <= 192 use the fused LN+1x1+dw3x3 kernels
    assert fused == 2 * (47 - 8 - 2) and kinds.count("dwconv") == 94 - fused
    assert kinds.count("mdta_gram") == 47 and kinds.count("prompt") == 3
    assert kinds.count("gemm") == 47 * 4 - fused + 3 + 3 + 5 + 3 + 1 and kinds[0] == "patch_embed"
    assert eng.kernels_per_forward() == len(eng.ops) + 47 + 3
    by, fl = map(sum, zip(*(op_cost(r) for r in eng.ops)))
    # 350 GFLOP and 7.29 GB per 256x256 image; 64x64 is 1/16 of that per image (+ weights)
    assert 0.9 < fl / (2 * 350.0e9 / 16) < 1.1
    with pytest.raises(RuntimeError):
        eng.run(torch.rand(2, 3, 64, 64))                 # no CPU execution path


def test_state_dict_roundtrip_and_refresh(model):
    torch.manual_seed(1)
    other = PromptIR(decoder=True).eval()
    missing = other.load_state_dict(model.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert [n for n, _ in other.named_parameters()] == [n for n, _ in model.named_parameters()]
    eng = Engine(other, 1, 32, 32, "cpu", torch.float32)
    x, _ = O.synthetic_batch(1, 32, 32, seed=9)
    y0 = emulator.run_program(eng, x)
    with torch.no_grad():
        other.output.weight.mul_(0.5)                     # in-place update bumps the version counter
    assert eng._current_version() != eng._param_version
    eng.refresh_weights()
    y1 = emulator.run_program(eng, x)
    ref = O.promptir_forward({k: v.detach() for k, v in other.state_dict().items()}, x)
    assert (y1 - ref).abs().max().item() < 2e-5 and (y0 - y1).abs().max().item() > 1e-4


def test_layernorm_fold_identity():
    """W.LN(x) == rstd*((W*g).x - mu*s) + W.b with s from the ROUNDED weights cancels the mean exactly."""
    torch.manual_seed(0)
    k, n = 96, 40
    x = torch.randn(50, k) * 0.3 + 5.0                    # large mean: the dangerous case
    w, g, b = torch.randn(n, k) / k ** 0.5, torch.rand(k) + 0.5, torch.randn(k)
    w16, s, t = packing.pack_pointwise(w, torch.bfloat16, gamma=g, beta=b)
    assert w16.shape == (n, 128) and float(w16[:, k:].abs().max()) == 0
    mu, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    folded = rstd * (x @ w16.float()[:, :k].t() - mu * s) + t
    direct = ((x - mu) * rstd) @ w16.float()[:, :k].t() + w @ b      # same rounded weights, LN applied first
    assert (folded - direct).abs().max().item() < 2e-4


def test_packing_layouts():
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = packing.pack_conv3x3(w, torch.float32)
    assert p.shape == (2, 9 * 64) and p[1, 4 * 64 + 2] == w[1, 2, 1, 1] and p[0, 63] == 0
    dw = torch.arange(4 * 9, dtype=torch.float32).reshape(4, 1, 3, 3)
    q = packing.pack_depthwise(dw, torch.float32)
    assert q.shape == (9, 4) and q[5, 3] == dw[3, 0, 1, 2]
    hp, gmap = packing.gdfn_maps(127, "cpu")
    assert hp == 128 and gmap[126] == 126 and gmap[127] == 128 and gmap.numel() == 254
    assert [packing.round_up(int(c * 2.66), 8) for c in (48, 96, 192, 384, 160, 320, 704)] == [128, 256, 512, 1024, 432, 856, 1872]


def test_module_errors_are_loud(model):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.rand(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        PromptIR()(torch.rand(1, 3, 32, 32))              # decoder=False: the reference crashes too

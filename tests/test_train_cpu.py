"""CPU: the training program (promptir_b200/train_engine.py) -- the hand-derived backward and its wiring -- interpreted by the torch
emulator in fp32 and compared with autograd of the oracle (the reference computes these gradients with autograd, train.py:37-46)."""
import pytest
import torch

import emulator
from oracle import promptir_oracle as O
from promptir_b200 import PromptIR
from promptir_b200.train_engine import TrainEngine


def oracle_grads(model, x, d_out, arch=O.ArchSpec()):
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xin = x.clone().requires_grad_(True)
    out = O.promptir_forward(sd, xin, arch)
    out.backward(d_out)
    return out.detach(), {k: v.grad for k, v in sd.items()}, xin.grad


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    m = PromptIR(decoder=True)
    with torch.no_grad():                      # make every parameter matter (temperature 1, LN affine 1/0 at init hide mistakes)
        for n, p in m.named_parameters():
            if n.endswith("temperature"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("bias"):
                p.copy_(torch.randn_like(p) * 0.2)
    return m


@pytest.mark.parametrize("shape", [(2, 32, 32), (1, 40, 24)])
def test_backward_program_matches_autograd(model, shape):
    B, H, W = shape
    x, _ = O.synthetic_batch(B, H, W, seed=3)
    torch.manual_seed(5)
    d_out = torch.randn(B, 3, H, W) / (B * 3 * H * W)
    ref_out, ref, ref_dx = oracle_grads(model, x, d_out)
    eng = TrainEngine(model, B, H, W, "cpu", torch.float32, input_grad=True)
    out, grads = emulator.run_train(eng, x, d_out)
    assert (out - ref_out).abs().max().item() < 2e-5
    dead = [n for n, g in ref.items() if g is None]
    assert len(dead) == 6                       # chnl_reduce1-3, reduce_noise_channel_1-3
    worst = max(((rel_err(grads[n], g), n) for n, g in ref.items() if g is not None))
    assert worst[0] < 2e-4, worst
    for n in dead:
        assert grads[n].abs().max().item() == 0.0
    assert rel_err(eng.d_img, ref_dx) < 2e-4


def test_backward_program_bias_and_biasfree():
    torch.manual_seed(2)
    m = PromptIR(decoder=True, bias=True, LayerNorm_type="BiasFree")
    x, _ = O.synthetic_batch(1, 32, 32, seed=4)
    d_out = torch.randn(1, 3, 32, 32) / 3072
    arch = O.ArchSpec(layernorm_type="BiasFree")
    ref_out, ref, _ = oracle_grads(m, x, d_out, arch)
    eng = TrainEngine(m, 1, 32, 32, "cpu", torch.float32)
    out, grads = emulator.run_train(eng, x, d_out)
    assert (out - ref_out).abs().max().item() < 2e-5
    worst = max(((rel_err(grads[n], g), n) for n, g in ref.items() if g is not None))
    assert worst[0] < 2e-4, worst


def test_loss_scale_is_undone():
    torch.manual_seed(0)
    m = PromptIR(decoder=True)
    x, _ = O.synthetic_batch(1, 32, 32, seed=4)
    d_out = torch.randn(1, 3, 32, 32) / 3072
    e1 = TrainEngine(m, 1, 32, 32, "cpu", torch.float32, grad_scale=1.0)
    e2 = TrainEngine(m, 1, 32, 32, "cpu", torch.float32, grad_scale=1024.0)
    _, g1 = emulator.run_train(e1, x, d_out)
    _, g2 = emulator.run_train(e2, x, d_out)
    assert max(rel_err(g2[n], g1[n]) for n in g1 if g1[n].abs().max() > 0) < 1e-5

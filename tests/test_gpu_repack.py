"""GPU: pir_repack (one launch) reproduces every derived weight cache that promptir_b200/packing.py states in torch -- for the
inference engine and the training engine (transposed / tap-flipped / GDFN-padded variants), in both storage types, and after the
parameters change (optimizer step, load_state_dict)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _check(pk):
    assert pk.jobs() > 0 and pk.checks
    kinds = set()
    for what, got, ref in pk.checks:
        kinds.add(what)
        assert got.shape == ref.shape and got.dtype == ref.dtype, what
        if got.dtype in (torch.float16, torch.bfloat16):
            assert torch.equal(got, ref), f"{what}: 16-bit cache differs from packing.py"
        else:                                           # fp32 vectors: same terms, different summation order
            torch.testing.assert_close(got, ref, rtol=2e-6, atol=2e-6, msg=what)
    return kinds


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("variant", ["default", "bias_biasfree"])
def test_repack_matches_packing_py(dt, variant, monkeypatch):
    from promptir_b200 import PromptIR
    from promptir_b200.engine import Engine
    from promptir_b200.train_engine import TrainEngine
    monkeypatch.setenv("PROMPTIR_B200_PACK_VERIFY", "1")
    torch.manual_seed(0)
    kw = dict(bias=True, LayerNorm_type="BiasFree") if variant == "bias_biasfree" else {}
    m = PromptIR(decoder=True, **kw).to(DEV)
    with torch.no_grad():                               # non-trivial LayerNorm affine, temperatures and biases
        for n, p in m.named_parameters():
            if "norm" in n or "temperature" in n or n.endswith(".bias"):
                p.add_(torch.randn_like(p) * 0.3)
    eng = Engine(m, 1, 32, 32, DEV, dt)
    torch.cuda.synchronize()
    kinds = _check(eng.pk)
    assert {"pointwise", "conv3x3", "depthwise", "prompt"} <= kinds
    teng = TrainEngine(m, 1, 32, 32, DEV, dt)
    torch.cuda.synchronize()
    _check(teng.pk)
    assert teng.pk.jobs() > eng.pk.jobs()


def test_refresh_after_parameter_update_is_one_launch():
    from promptir_b200 import PromptIR, _lib
    torch.manual_seed(0)
    m = PromptIR(decoder=True).eval().to(DEV)
    x = torch.rand(1, 3, 32, 32, device=DEV)
    with torch.no_grad():
        y0 = m(x).clone()
        for p in m.parameters():
            p.mul_(1.01)                                # what optimizer.step() does: in-place update, version bump
        n0 = _lib.launch_count
        y1 = m(x).clone()
        launches = _lib.launch_count - n0
        m2 = PromptIR(decoder=True).eval().to(DEV)
        m2.load_state_dict(m.state_dict())
        y2 = m2(x)
    # the forward itself is a CUDA-graph replay (not counted by the binding): the only launch issued is pir_repack
    assert launches == 1, "exactly one launch (pir_repack) refreshes the caches"
    assert not torch.equal(y0, y1)
    assert torch.equal(y1, y2), "refreshed caches == caches of a freshly built engine with the same parameters"


def test_engine_is_rebuilt_when_parameter_storage_moves():
    from promptir_b200 import PromptIR
    torch.manual_seed(0)
    m = PromptIR(decoder=True).eval().to(DEV)
    x = torch.rand(1, 3, 32, 32, device=DEV)
    with torch.no_grad():
        m(x)
        e0 = m.engine_for(1, 32, 32, torch.device(DEV))
        sd = {k: (v * 0.5).clone() for k, v in m.state_dict().items()}
        m.load_state_dict(sd, assign=True)              # replaces the parameter tensors: raw pointers in the old engine are stale
        y = m(x)
        e1 = m.engine_for(1, 32, 32, torch.device(DEV))
        m2 = PromptIR(decoder=True).eval().to(DEV)
        m2.load_state_dict(sd)
        assert e1 is not e0
        assert torch.equal(y, m2(x))

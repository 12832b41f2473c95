"""GPU parity of PromptXRestormer (BASELINE.json configs[4]) through the C ABI: the OCAB kernel against its
specification, the XEngine program op by op, and the whole forward against the golden outputs of the real reference.
Tolerances, stated here (the north star fixes 2e-3 / 0.02 dB for PromptIR only): this network has 200 residual sub-layers against
PromptIR's 94 (four per block, plus three 160/320/704-wide prompt blocks), so 16-bit rounding accumulates further:
fp16 max-abs <= 4e-3 on clamp(out,0,1) vs the fp32 reference (measured 2.2e-3..2.8e-3), bf16 <= 3e-2 (measured 1.8e-2..1.9e-2),
dPSNR <= 0.02 dB for both (measured <= 0.0023 dB).  The limits are justified by the 16-bit-reference test below: the reference
network itself, executed by PyTorch eager in fp16 / bf16 on the same GPU, is 4.3e-3..1.8e-2 / 1.8e-2..2.2e-2 from its fp32 result, and
this build has to be at least that close (test_16bit_build_is_as_close_to_fp32_as_the_reference_run_in_16bit)."""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch

import emulator
from oracle import promptir_oracle as O
from promptir_b200 import PromptXRestormer, ops
from promptir_b200.xengine import XEngine

pytestmark = pytest.mark.gpu
DEV = "cuda"
MAXABS = {torch.float16: 4e-3, torch.bfloat16: 3e-2}


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return PromptXRestormer().eval().to(DEV)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("heads,shape", [(2, (2, 16, 24)), (3, (1, 8, 8)), (8, (1, 32, 16))])
def test_ocab_kernel(dt, heads, shape):
    """Windows on every border (zero-padded keys take part in the softmax), channel-slice views, several head counts."""
    B, H, W = shape
    inner = 16 * heads
    torch.manual_seed(heads)
    buf = (torch.randn(B, H, W, 3 * inner + 8, device=DEV) * 1.5).to(dt)
    qkv = buf[..., 8:]
    rel_h = torch.randn(23, 16, device=DEV) * 0.25
    rel_w = torch.randn(23, 16, device=DEV) * 0.25
    obuf = torch.zeros(B, H, W, inner + 16, device=DEV, dtype=dt)
    out = obuf[..., 16:]
    ops.ocab(qkv, rel_h, rel_w, out, heads=heads)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = torch.zeros_like(out)
    emulator.emu_ocab(dict(qkv=qkv, rel_h=rel_h, rel_w=rel_w, out=ref, heads=heads))
    err = (out.float() - ref.float()).abs()
    lim = (2e-3 if dt == torch.bfloat16 else 3e-4) * ref.float().abs().max().item() + (2 ** -7 if dt == torch.bfloat16 else 2 ** -10) * ref.float().abs()
    assert not (err > lim).any(), f"max err {err.max().item():.4g} ref max {ref.float().abs().max().item():.4g}"
    assert obuf[..., :16].abs().max().item() == 0


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_xengine_op_by_op(model, golden_dir, dt):
    g = np.load(os.path.join(golden_dir, "xrestormer_seed0.npz"))
    x = torch.from_numpy(g["x64x128_in"]).to(DEV)
    eng = XEngine(model, x.shape[0], x.shape[2], x.shape[3], DEV, dt)
    eng.img_in.copy_(x)
    s = torch.cuda.current_stream().cuda_stream
    failures, i, ops_ = [], 0, eng.ops
    while i < len(ops_) and len(failures) < 12:
        r = ops_[i]
        if r["kind"] == "mdta_gram":
            fin = ops_[i + 1]
            r["launch"](s)
            fin["launch"](s)
            torch.cuda.synchronize()
            got = fin["wfold"].clone()
            emulator.emu_mdta_finalize(fin)
            ref, name, i = fin["wfold"], f"{i}:mdta(C={fin['wfold'].shape[1]})", i + 2
        else:
            out = r["out"]
            before = out.clone()
            r["launch"](s)
            torch.cuda.synchronize()
            got = out.clone()
            out.copy_(before)
            emulator.DISPATCH[r["kind"]](r)
            ref, name, i = out, f"{i}:{r['kind']}:{r.get('tag', '')}{tuple(out.shape)}", i + 1
        err = (got.float() - ref.float()).abs()
        lim = (2e-3 if dt == torch.bfloat16 else 3e-4) * max(1.0, ref.float().abs().max().item()) + (2 ** -7 if dt == torch.bfloat16 else 2 ** -10) * ref.float().abs()
        if r["kind"] == "gemm" and r["out_mode"] == 3:
            lim = torch.full_like(err, 1e-4)
        if r["kind"] == "pwdw":                            # nine taps accumulated in fp16: up to 9 * 2^-11 of the largest partial sum
            lim = lim + 3e-3 * ref.float().abs().max().item()
        nbad = int((err > lim).sum())
        if nbad or torch.isnan(got.float()).any():
            failures.append(f"{name}: {nbad}/{err.numel()} bad, max err {err.max().item():.4g}, ref max {ref.float().abs().max().item():.4g}")
    assert not failures, "\n".join(failures)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case,seed", [("x64", 1), ("x64x128", 2), ("x128", 3)])
def test_forward_matches_reference_golden(model, golden_dir, dt, case, seed):
    g = np.load(os.path.join(golden_dir, "xrestormer_seed0.npz"))
    x = torch.from_numpy(g[case + "_in"]).to(DEV)
    yref = torch.from_numpy(g[case + "_out"]).to(DEV)
    model.compute_dtype = dt
    with torch.no_grad():
        y_eager = model.engine_for(x.shape[0], x.shape[2], x.shape[3], x.device).run(x, use_graph=False)
        y = model(x)
    assert torch.equal(y, y_eager), "graph replay differs from eager launches"
    err = (y.clamp(0, 1) - yref.clamp(0, 1)).abs().max().item()
    _, clean = O.synthetic_batch(x.shape[0], x.shape[2], x.shape[3], seed=seed)
    dpsnr = abs(O.psnr(y.cpu(), clean) - O.psnr(yref.cpu(), clean))
    print(f"[x parity] {case} {dt}: max-abs(clamped) {err:.3e} raw {(y - yref).abs().max().item():.3e} dPSNR {dpsnr:.4f} dB")
    assert err <= MAXABS[dt] and dpsnr <= 0.02


def _errors_vs_16bit_reference(model, dt, x):
    """-> (max-abs, rms) of this build and of the reference algorithm executed with `dt` tensors by PyTorch eager on the same GPU,
    both against the fp32 run of the reference algorithm (clamped max-abs, raw rms)."""
    from oracle import xrestormer_oracle as XO
    sd32 = {k: v.detach().to(DEV) for k, v in model.state_dict().items()}
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ref32 = XO.xrestormer_forward(sd32, x.to(DEV)).float()
            ref16 = XO.xrestormer_forward({k: v.to(dt) for k, v in sd32.items()}, x.to(DEV).to(dt)).float()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    model.compute_dtype = dt
    with torch.no_grad():
        got = model(x.to(DEV)).float()
    ours = float((got.clamp(0, 1) - ref32.clamp(0, 1)).abs().max()), float((got - ref32).pow(2).mean().sqrt())
    eager = float((ref16.clamp(0, 1) - ref32.clamp(0, 1)).abs().max()), float((ref16 - ref32).pow(2).mean().sqrt())
    return ours, eager


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_16bit_build_is_as_close_to_fp32_as_the_reference_run_in_16bit(model, dt):
    """Why the limits above are 4e-3 / 3e-2 and not PromptIR's 2e-3 (the fair oracle for a 16-bit build is the
    reference executed in that type on the same GPU): the reference network itself, run by PyTorch eager with fp16 / bf16 tensors,
    lands this far from its own fp32 result.  This build (16-bit storage, fp32 accumulation, fp32 LayerNorm / softmax statistics) must
    not be further away than that, in rms over the image (5 % slack for the eager side's algorithm choices) and in max-abs
    (single-pixel statistic: 25 % slack).  Measured on B200: fp16 this build 2.3e-3 / 2.6e-3 max-abs (rms 5.6e-4 / 5.9e-4) against
    1.8e-2 / 4.3e-3 (rms 3.6e-3 / 1.0e-3) for the reference in fp16 eager; bf16 1.6e-2 / 1.9e-2 (rms 4.7e-3 / 4.8e-3) against
    1.8e-2 / 2.2e-2 (rms 5.1e-3 / 5.2e-3)."""
    for shape, seed in (((2, 64, 64), 5), ((1, 128, 128), 6)):
        x, _ = O.synthetic_batch(*shape, seed=seed)
        (err_ours, rms_ours), (err_eager, rms_eager) = _errors_vs_16bit_reference(model, dt, x)
        print(f"[x 16-bit] {dt} {shape}: this build max-abs {err_ours:.3e} rms {rms_ours:.3e}; reference in {dt} eager max-abs {err_eager:.3e} rms {rms_eager:.3e}")
        assert rms_ours <= 1.05 * rms_eager and err_ours <= 1.25 * err_eager


def test_errors_are_loud(model):
    with pytest.raises(RuntimeError):
        model(torch.rand(1, 3, 96, 64, device=DEV))           # not a multiple of 64


def _small_x_model(seed=0):
    torch.manual_seed(seed)
    m = PromptXRestormer(num_blocks=[1, 1, 1, 2], num_refinement_blocks=1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("temperature"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("bias"):
                p.copy_(torch.randn_like(p) * 0.2)
    return m


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_xtrain_engine_op_by_op(dt):
    """Every launch of the PromptXRestormer training programs (incl. pir_ocab_bwd) against its torch restatement."""
    from test_gpu_train import _run_ops
    from promptir_b200.xtrain_engine import XTrainEngine
    m = _small_x_model().to(DEV)
    B, H, W = 2, 64, 128
    x, _ = O.synthetic_batch(B, H, W, seed=3)
    eng = XTrainEngine(m, B, H, W, DEV, dt, grad_scale=1.0 if dt == torch.bfloat16 else 64.0, input_grad=dt == torch.bfloat16)
    eng.img_in.copy_(x.to(DEV))
    s = torch.cuda.current_stream().cuda_stream
    failures = []
    _run_ops(eng.fwd_ops, s, failures, dt)
    assert not failures, "forward:\n" + "\n".join(failures)
    torch.manual_seed(5)
    eng.d_out.copy_(torch.randn(B, 3, H, W, device=DEV) / 64)
    _run_ops(eng.bwd_ops, s, failures, dt)
    assert not failures, "backward:\n" + "\n".join(failures)


@pytest.mark.parametrize("dt,per_lim,cos_lim", [(torch.bfloat16, 1e-1, 0.998), (torch.float16, 3e-2, 0.9999)])
def test_x_loss_backward_matches_oracle_autograd(dt, per_lim, cos_lim):
    """loss.backward() through the drop-in PromptXRestormer vs fp32 autograd of the CPU oracle (stated limits: per-tensor relative
    L2 error 10 % / 3 %, flat-gradient cosine 0.998 / 0.9999 for bf16 / fp16)."""
    from oracle import xrestormer_oracle as XO
    m = _small_x_model(seed=1).to(DEV).train()
    m.compute_dtype = dt
    B, H, W = 1, 64, 64
    x, clean = O.synthetic_batch(B, H, W, seed=7)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    ref_loss = torch.nn.functional.l1_loss(XO.xrestormer_forward(sd, x, num_blocks=(1, 1, 1, 2), num_refinement_blocks=1), clean)
    ref_loss.backward()
    loss = torch.nn.functional.l1_loss(m(x.to(DEV)), clean.to(DEV))
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-2 * abs(ref_loss.item()) + 1e-4
    fg, fr, worst = [], [], (0.0, "")
    gmax = max(v.grad.norm().item() for v in sd.values())
    for n, p in m.named_parameters():
        g, r = p.grad.detach().cpu().float(), sd[n].grad
        assert torch.isfinite(g).all(), n
        if r.norm().item() > 1e-3 * gmax:
            worst = max(worst, (((g - r).norm() / r.norm()).item(), n))
        fg.append(g.reshape(-1))
        fr.append(r.reshape(-1))
    fg, fr = torch.cat(fg), torch.cat(fr)
    cos = torch.nn.functional.cosine_similarity(fg, fr, dim=0).item()
    print(f"[x train parity] {dt}: loss {loss.item():.5f} (oracle {ref_loss.item():.5f}) flat-grad cos {cos:.6f} "
          f"rel-L2 {((fg - fr).norm() / fr.norm()).item():.4f} worst tensor {worst[1]} {worst[0]:.4f}")
    assert cos >= cos_lim and worst[0] <= per_lim, (cos, worst)


def test_x_bf16_gradients_are_as_close_to_fp32_as_the_reference_trained_in_bf16():
    """The 10 % per-tensor limit of the test above against what autograd of the reference algorithm delivers in bf16 on the same GPU
    (all-bf16 tensors, and torch.autocast(bfloat16) with fp32 parameters), all three compared with fp32 autograd of the CPU oracle:
    this build's bf16 training path has to be in the same range -- flat-gradient relative L2 error within 1.5x, worst tensor within
    2x of the reference's own bf16 autograd (on this small network the all-bf16 run is already within a few per cent of fp32)."""
    from oracle import xrestormer_oracle as XO
    kw = dict(num_blocks=(1, 1, 1, 2), num_refinement_blocks=1)
    m = _small_x_model(seed=1).to(DEV).train()
    m.compute_dtype = torch.bfloat16
    x, clean = O.synthetic_batch(1, 64, 64, seed=7)
    xd, cd = x.to(DEV), clean.to(DEV)
    l1 = torch.nn.functional.l1_loss
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    l1(XO.xrestormer_forward(sd, x, **kw), clean).backward()
    ref = {k: v.grad for k, v in sd.items()}
    gmax = max(r.norm().item() for r in ref.values())

    def errors(grads):
        fg, fr, worst = [], [], 0.0
        for n, g in grads.items():
            g, r = g.detach().cpu().float(), ref[n]
            if r.norm().item() > 1e-3 * gmax:
                worst = max(worst, ((g - r).norm() / r.norm()).item())
            fg.append(g.reshape(-1))
            fr.append(r.reshape(-1))
        fg, fr = torch.cat(fg), torch.cat(fr)
        return ((fg - fr).norm() / fr.norm()).item(), worst

    l1(m(xd), cd).backward()
    ours = errors({n: p.grad for n, p in m.named_parameters()})
    sd16 = {k: v.detach().to(torch.bfloat16).requires_grad_(True) for k, v in m.state_dict().items()}
    l1(XO.xrestormer_forward(sd16, xd.bfloat16(), **kw).float(), cd).backward()
    eager16 = errors({k: v.grad for k, v in sd16.items()})
    sd32 = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = XO.xrestormer_forward(sd32, xd, **kw)
    l1(out.float(), cd).backward()
    autocast = errors({k: v.grad for k, v in sd32.items()})
    for name, (rel, worst) in (("this build bf16", ours), ("reference all-bf16 autograd", eager16), ("reference bf16 autocast", autocast)):
        print(f"[x train 16-bit] {name}: flat-grad rel-L2 {rel:.4f} worst tensor {worst:.4f}")
    # measured on B200: this build 0.44 % flat / 4.5 % worst tensor; all-bf16 autograd 0.74 % / 3.6 %; bf16 autocast 0.48 % / 2.4 %
    assert ours[0] <= 1.5 * max(eager16[0], autocast[0]) and ours[1] <= 2.0 * max(eager16[1], autocast[1])

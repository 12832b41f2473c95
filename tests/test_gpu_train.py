"""GPU parity of the training path (forward that keeps activations + hand-written backward), through the C ABI.

  * op by op: every launch of the TrainEngine program against the torch restatement of the same op (tests/emulator.py) on the
    same device buffers -- the emulator's result flows on, so errors do not compound and the first bad kernel is named;
  * end to end: parameter gradients of `loss.backward()` through the drop-in module against fp32 autograd of the CPU oracle
    (that is how the reference computes them: train.py:37-46).  Tolerance (stated here; the north star only fixes the forward
    tolerance): per-parameter relative L2 error <= 8e-2 (bf16) / 2e-2 (fp16) and cosine of the whole flat gradient >= 0.999 /
    0.9999 -- 16-bit activations and activation gradients through 47 residual blocks, fp32 accumulation everywhere.
"""
from __future__ import annotations

import pytest
import torch

import emulator
from oracle import promptir_oracle as O
from promptir_b200 import PromptIR
from promptir_b200.train_engine import TrainEngine

pytestmark = pytest.mark.gpu
DEV = "cuda"


def perturbed_model(seed=0, **kw):
    torch.manual_seed(seed)
    m = PromptIR(decoder=True, **kw)
    with torch.no_grad():                      # make every parameter matter (temperature 1, LN affine 1/0 at init hide mistakes)
        for n, p in m.named_parameters():
            if n.endswith("temperature"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand_like(p) + 0.5)
            elif "norm" in n and n.endswith("bias"):
                p.copy_(torch.randn_like(p) * 0.2)
    return m


OUTPUTS = {"ln_fwd": ["xhat", "rstd"], "ln_bwd": ["g"], "gemm": ["out"], "dwconv": ["out"], "patch_embed": ["out"], "gate_bwd": ["y"],
           "shuffle": ["out"], "bcast_add": ["g"], "to_nhwc16": ["out"], "wgrad_fin": ["dst_w", "dst_gamma", "dst_beta", "dst_bias"],
           "dw_wgrad": ["dst_w", "dst_bias"], "mdta_bwd": ["dst_wo", "dst_temp", "dst_bias", "wqk", "wft"],
           "prompt_bwd": ["dst_prompt", "dst_lin_w", "dst_lin_b", "demb"], "prompt": ["out", "weights_out"], "ocab": ["out"],
           "ocab_bwd": ["dqkv", "dst_rel_h", "dst_rel_w"]}


def _limit(ref, dt16):
    ref = ref.float()
    mx = max(ref.abs().max().item(), 1e-30)
    if ref.dtype == torch.float32 and not dt16:
        return 2e-3 * mx
    atol, rtol = (2e-3, 2 ** -7)
    return atol * mx + rtol * ref.abs() + 1.2e-7          # + one fp16 subnormal step (6e-8): small fp16 gradients are quantised


def _compare(name, got, ref, failures, loose=1.0):
    is16 = ref.dtype != torch.float32
    err = (got.float() - ref.float()).abs()
    lim = _limit(ref, is16) * loose
    nbad = int((err > lim).sum())
    if nbad or torch.isnan(got.float()).any():
        failures.append(f"{name}: {nbad}/{err.numel()} bad, max err {err.max().item():.4g}, ref max {ref.float().abs().max().item():.4g}")


def _run_ops(ops_, s, failures, dt, limit=16):
    i = 0
    while i < len(ops_) and len(failures) < limit:
        r = ops_[i]
        kind = r["kind"]
        tag = f"{i}:{kind}:{r.get('tag', '')}"
        if kind == "mdta_gram":
            fin = ops_[i + 1]
            r["launch"](s)
            fin["launch"](s)
            torch.cuda.synchronize()
            got = fin["wfold"].clone()
            emulator.emu_mdta_finalize(fin)
            _compare(tag + f"(C={fin['wfold'].shape[1]})", got, fin["wfold"], failures)
            i += 2
            continue
        if kind == "wgrad":
            P, taps, M, N = r["P"], r["taps"], r["M"], r["N"]
            r["launch"](s)
            torch.cuda.synchronize()
            part, cs = emulator._wg_views(r["ws"], r)
            groups = r["a"].shape[0] if r["per_image"] else 1
            got = part.view(groups, -1, taps, M, N).sum(1).clone()
            got_cs = cs.view(groups, -1, M).sum(1).clone() if cs is not None else None
            emulator.emu_wgrad(r)
            ref = part.view(groups, -1, taps, M, N).sum(1)
            scale = (r["a"].float().norm() * r["b"].float().norm()).item() / max(M * N, 1) ** 0.5
            err = (got - ref).abs().max().item()
            if not err <= 1e-3 * max(ref.abs().max().item(), 1e-30) + 1e-6 * scale:
                failures.append(f"{tag} M={M} N={N} taps={taps} P={P}: max err {err:.4g}, ref max {ref.abs().max().item():.4g}")
            if cs is not None:
                ref_cs = cs.view(groups, -1, M).sum(1)
                e2 = (got_cs - ref_cs).abs().max().item()
                if not e2 <= 1e-3 * max(ref_cs.abs().max().item(), 1e-30) + 1e-6 * r["a"].float().abs().sum().item() / M:
                    failures.append(f"{tag} colsum: max err {e2:.4g}, ref max {ref_cs.abs().max().item():.4g}")
            i += 1
            continue
        outs = [(k, r[k]) for k in OUTPUTS[kind] if r.get(k) is not None]
        before = [t.clone() for _, t in outs]
        ws_before = r["ws"].clone() if kind == "wgrad_fin" else None       # the finalize consumes (reduces in place) its partials
        r["launch"](s)
        torch.cuda.synchronize()
        got = [t.clone() for _, t in outs]
        for (_, t), b in zip(outs, before):
            t.copy_(b)
        if ws_before is not None:
            r["ws"].copy_(ws_before)
        (emulator.emu_prompt_train if kind == "prompt" else emulator.DISPATCH[kind])(r)
        for (k, t), g in zip(outs, got):
            loose = 4.0 if kind in ("mdta_bwd", "prompt_bwd", "ocab_bwd") else 1.0
            if kind == "gemm" and r["out_mode"] == 3:
                if (g - t).abs().max().item() > 1e-4 * max(1.0, t.abs().max().item()):
                    failures.append(f"{tag}.{k}: max err {(g - t).abs().max().item():.4g}")
                continue
            _compare(f"{tag}.{k}{tuple(t.shape)}", g, t, failures, loose)
        i += 1


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_train_engine_op_by_op(dt):
    m = perturbed_model().to(DEV)
    B, H, W = 2, 40, 24
    x, _ = O.synthetic_batch(B, H, W, seed=3)
    eng = TrainEngine(m, B, H, W, DEV, dt, grad_scale=1.0 if dt == torch.bfloat16 else 64.0, input_grad=dt == torch.bfloat16)
    eng.img_in.copy_(x.to(DEV))
    s = torch.cuda.current_stream().cuda_stream
    failures = []
    _run_ops(eng.fwd_ops, s, failures, dt)
    assert not failures, "forward:\n" + "\n".join(failures)
    torch.manual_seed(5)
    eng.d_out.copy_(torch.randn(B, 3, H, W, device=DEV) / 64)
    _run_ops(eng.bwd_ops, s, failures, dt)
    assert not failures, "backward:\n" + "\n".join(failures)


def test_train_engine_op_by_op_bias_biasfree():
    m = perturbed_model(seed=2, bias=True, LayerNorm_type="BiasFree").to(DEV)
    B, H, W = 1, 32, 32
    x, _ = O.synthetic_batch(B, H, W, seed=4)
    eng = TrainEngine(m, B, H, W, DEV, torch.bfloat16)
    eng.img_in.copy_(x.to(DEV))
    s = torch.cuda.current_stream().cuda_stream
    failures = []
    _run_ops(eng.fwd_ops, s, failures, torch.bfloat16)
    eng.d_out.copy_(torch.randn(B, 3, H, W, device=DEV) / 64)
    _run_ops(eng.bwd_ops, s, failures, torch.bfloat16)
    assert not failures, "\n".join(failures)


def _oracle_grads(model, x, target):
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    out = O.promptir_forward(sd, x)
    loss = torch.nn.functional.l1_loss(out, target)            # train.py:28,43
    loss.backward()
    return loss.item(), {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("dt,per_lim,cos_lim", [(torch.bfloat16, 8e-2, 0.999), (torch.float16, 2e-2, 0.9999)])
def test_loss_backward_matches_oracle_autograd(dt, per_lim, cos_lim):
    """The call train.py makes: restored = net(x); loss = L1(restored, clean); loss.backward()."""
    m = perturbed_model(seed=1).to(DEV).train()
    m.compute_dtype = dt
    B, H, W = 2, 64, 64
    x, clean = O.synthetic_batch(B, H, W, seed=7)
    ref_loss, ref = _oracle_grads(m, x, clean)
    for use_graph in (False, True):
        m.use_cuda_graph = use_graph
        m.zero_grad(set_to_none=True)
        out = m(x.to(DEV))
        loss = torch.nn.functional.l1_loss(out, clean.to(DEV))
        loss.backward()
        assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss) + 1e-4
        flat_g, flat_r, worst = [], [], (0.0, "")
        for n, p in m.named_parameters():
            if ref[n] is None:
                assert p.grad is None, n                     # dead parameters get no gradient, like autograd
                continue
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
            g, r = p.grad.detach().cpu().float(), ref[n]
            rel = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
            if r.norm().item() > 1e-3 * max(t.norm().item() for t in ref.values() if t is not None):    # ignore negligible tensors
                worst = max(worst, (rel, n))
            flat_g.append(g.reshape(-1))
            flat_r.append(r.reshape(-1))
        fg, fr = torch.cat(flat_g), torch.cat(flat_r)
        cos = torch.nn.functional.cosine_similarity(fg, fr, dim=0).item()
        rel_all = ((fg - fr).norm() / fr.norm()).item()
        print(f"[train parity] {dt} graph={use_graph}: loss {loss.item():.5f} (oracle {ref_loss:.5f}) flat-grad cos {cos:.6f} "
              f"rel-L2 {rel_all:.4f} worst tensor {worst[1]} {worst[0]:.4f}")
        assert cos >= cos_lim and worst[0] <= per_lim, (cos, worst)


def _grad_errors(grads, ref):
    """flat relative L2 error, cosine and the worst non-negligible tensor of {name: gradient} against the fp32 reference gradients."""
    big = 1e-3 * max(t.norm().item() for t in ref.values() if t is not None)
    flat_g, flat_r, worst = [], [], (0.0, "")
    for n, r in ref.items():
        if r is None or grads.get(n) is None:
            continue
        g = grads[n].detach().cpu().float()
        if r.norm().item() > big:
            worst = max(worst, (((g - r).norm() / r.norm()).item(), n))
        flat_g.append(g.reshape(-1))
        flat_r.append(r.reshape(-1))
    fg, fr = torch.cat(flat_g), torch.cat(flat_r)
    return ((fg - fr).norm() / fr.norm()).item(), torch.nn.functional.cosine_similarity(fg, fr, dim=0).item(), worst


def test_bf16_gradients_are_as_close_to_fp32_as_the_reference_trained_in_bf16():
    """The limits of the test above (worst tensor 8 %) against what the reference itself delivers when its training step runs in bf16
    on the same GPU: autograd of the reference algorithm with bf16 parameters and activations, and under
    torch.autocast(bfloat16) with fp32 parameters, both compared with fp32 autograd.  This library's bf16 training path (bf16 storage,
    fp32 accumulation, fp32 weight gradients) must not be further from fp32 than either of them."""
    m = perturbed_model(seed=1).to(DEV).train()
    m.compute_dtype = torch.bfloat16
    x, clean = O.synthetic_batch(2, 64, 64, seed=7)
    _, ref = _oracle_grads(m, x, clean)
    xd, cd = x.to(DEV), clean.to(DEV)
    m.zero_grad(set_to_none=True)
    torch.nn.functional.l1_loss(m(xd), cd).backward()
    ours = _grad_errors({n: p.grad for n, p in m.named_parameters()}, ref)
    sd16 = {k: v.detach().to(torch.bfloat16).requires_grad_(True) for k, v in m.state_dict().items()}
    torch.nn.functional.l1_loss(O.promptir_forward(sd16, xd.bfloat16()).float(), cd).backward()
    eager16 = _grad_errors({k: v.grad for k, v in sd16.items()}, ref)
    sd32 = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.promptir_forward(sd32, xd)
    torch.nn.functional.l1_loss(out.float(), cd).backward()
    autocast = _grad_errors({k: v.grad for k, v in sd32.items()}, ref)
    for name, (rel, cos, worst) in (("this build bf16", ours), ("reference all-bf16 autograd", eager16), ("reference bf16 autocast", autocast)):
        print(f"[train 16-bit] {name}: flat-grad rel-L2 {rel:.4f} cos {cos:.6f} worst tensor {worst[1]} {worst[0]:.4f}")
    # measured on B200: this build rel-L2 0.53 % (worst tensor 4.7 %), all-bf16 autograd 1.34 % (6.2 %), bf16 autocast 0.99 % (7.3 %)
    assert ours[0] <= eager16[0] and ours[0] <= autocast[0] and ours[2][0] <= 1.25 * eager16[2][0]


def test_optimizer_step_refreshes_packed_weights():
    """AdamW updates the fp32 parameters in place (train.py:52-56); the next forward (training or inference program) must see them."""
    m = perturbed_model(seed=1).to(DEV).train()
    x, clean = O.synthetic_batch(1, 32, 32, seed=9)
    xd, cd = x.to(DEV), clean.to(DEV)
    opt = torch.optim.AdamW(m.parameters(), lr=2e-4)
    out0 = None
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        out = m(xd)
        out0 = out.detach().clone() if out0 is None else out0
        torch.nn.functional.l1_loss(out, cd).backward()
        opt.step()
    for n in ("chnl_reduce1.weight", "reduce_noise_channel_2.weight"):
        assert dict(m.named_parameters())[n].grad is None
    ref = O.promptir_forward({k: v.detach().cpu() for k, v in m.state_dict().items()}, x)
    out_train = m(xd).detach()
    with torch.no_grad():
        out_infer = m(xd)
    assert (out_train.cpu() - ref).abs().max().item() <= 2e-2 and (out_infer.cpu() - ref).abs().max().item() <= 2e-2
    assert (out0.cpu() - ref).abs().max().item() > 1e-3              # the update was visible


def test_stale_backward_is_loud():
    m = perturbed_model(seed=1).to(DEV).train()
    x, _ = O.synthetic_batch(1, 32, 32, seed=9)
    x = x.to(DEV)
    out1 = m(x)
    out2 = m(x)
    out2.sum().backward()
    with pytest.raises(RuntimeError):
        out1.sum().backward()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_gate_bwd_kernel(dt):
    """pir_gate_bwd (the unfused gate backward, exact erf) on a channel-slice view."""
    from promptir_b200 import ops
    torch.manual_seed(0)
    B, H, W, Cc = 2, 9, 13, 40
    ybuf = (torch.randn(B, H, W, 2 * Cc + 8, device=DEV) * 1.5).to(dt)
    y = ybuf[..., 8:]
    dg = (torch.randn(B, H, W, Cc, device=DEV) * 0.3).to(dt)
    ref = y.clone()
    emulator.emu_gate_bwd(dict(y=ref, dgt=dg))
    ops.gate_bwd(y, dg)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    failures = []
    _compare("gate_bwd", y, ref, failures)
    assert not failures, failures


@pytest.mark.parametrize("shape", [(1, 8, 8), (3, 16, 40), (1, 8, 136)])
def test_training_extreme_shapes(shape):
    """Smallest legal images (latent = one pixel), thin strips, odd batches: gradients vs fp32 autograd of the oracle."""
    B, H, W = shape
    m = perturbed_model(seed=3).to(DEV).train()
    # 8 x 8: the level-4 attention normalises over ONE pixel, so 1 / (|q| |k|) is unbounded; with fp16's loss scale that leaves the fp16
    # range (saturated, i.e. clipped gradients) -- bf16, the training default, has the range
    m.compute_dtype = torch.bfloat16 if H * W == 64 else torch.float16
    x, clean = O.synthetic_batch(B, H, W, seed=H * W)
    ref_loss, ref = _oracle_grads(m, x, clean)
    loss = torch.nn.functional.l1_loss(m(x.to(DEV)), clean.to(DEV))
    loss.backward()
    fg = torch.cat([p.grad.reshape(-1).cpu() for n, p in m.named_parameters() if ref[n] is not None])
    fr = torch.cat([ref[n].reshape(-1) for n, p in m.named_parameters() if ref[n] is not None])
    assert torch.isfinite(fg).all()
    rel = ((fg - fr).norm() / fr.norm()).item()
    print(f"[train extreme] {shape}: loss {loss.item():.5f} vs {ref_loss:.5f}, flat-gradient rel-L2 {rel:.4f}")
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss) + 1e-4 and rel <= (8e-2 if H * W == 64 else 2e-2)


def test_fp16_backward_reports_loss_scale_overflow():
    """fp16 storage runs the backward under a static loss scale of 65536: an upstream gradient of order 1 (sum-reduced loss) leaves
    the 16-bit range.  The engine says so (TrainEngine.overflowed) instead of silently handing non-finite / clipped gradients to the
    optimizer; a mean-reduced loss does not trip it, and bf16 (scale 1) never does."""
    from promptir_b200 import PromptIR
    from promptir_b200.train_engine import TrainEngine
    torch.manual_seed(0)
    m = PromptIR(decoder=True).to("cuda").train()
    x = torch.rand(1, 3, 32, 32, device="cuda")
    eng = TrainEngine(m, 1, 32, 32, "cuda", torch.float16)
    eng.forward(x)
    eng.backward(torch.full((1, 3, 32, 32), 1.0 / (3 * 32 * 32), device="cuda"))      # d(mean loss)/d(out)
    assert not eng.overflowed() and bool(torch.isfinite(eng.grad_flat).all())
    eng.forward(x)
    eng.backward(torch.full((1, 3, 32, 32), 4.0, device="cuda"))                        # 4 * 65536 does not fit fp16
    assert eng.overflowed()
    eng16 = TrainEngine(m, 1, 32, 32, "cuda", torch.bfloat16)
    eng16.forward(x)
    eng16.backward(torch.full((1, 3, 32, 32), 4.0, device="cuda"))
    assert not eng16.overflowed()

"""Torch (CPU) interpreter of the engine's op records.  TEST INFRASTRUCTURE: it restates what each CUDA kernel
is specified to compute (include/promptir_b200.h) on the engine's own buffers and packed weights, so the
program wiring, the concat folding and the weight packing can be checked against the oracle without a GPU,
and the expected rounding error of the 16-bit pipeline can be estimated."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from promptir_b200._lib import (LN_BIASFREE, OUT_FINAL_NCHW32, OUT_NHWC16, OUT_NHWC32, OUT_SHUFFLE16, OUT_UNSHUFFLE16)


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1)


def emu_gemm(r):
    a, w, out, n, taps = r["a"], r["w"], r["out"], r["n"], r["taps"]
    B, H, W, K = a.shape
    kpad = (K + 63) // 64 * 64
    A = a.float()
    if r["w_batched"]:
        wt = w.float().view(B, -1, kpad)[:, :n, :K]                     # [B, n, K]
        acc = torch.einsum("bhwk,bnk->bhwn", A, wt)
    else:
        wt = w.float()[:n].view(n, taps, kpad)[:, :, :K]                # [n, taps, K]
        if taps == 1:
            acc = torch.einsum("bhwk,nk->bhwn", A, wt[:, 0])
        else:
            cw = wt.view(n, 3, 3, K).permute(0, 3, 1, 2).contiguous()
            acc = _nhwc(F.conv2d(_nchw(a), cw, padding=1))
    if r["ln_mode"]:
        mu = A.mean(-1, keepdim=True)
        var = ((A * A).mean(-1, keepdim=True) - mu * mu).clamp_min(0)
        rstd = torch.rsqrt(var + 1e-5)
        if r["ln_mode"] == LN_BIASFREE:
            mu = torch.zeros_like(mu)
        acc = rstd * (acc - mu * r["ln_s"][:n].view(1, 1, 1, -1))
    if r["vec_t"] is not None:
        acc = acc + r["vec_t"][:n].view(1, 1, 1, -1)
    if r["res"] is not None:
        acc = acc + r["res"].float()
    mode = r["out_mode"]
    if mode in (OUT_NHWC16, OUT_NHWC32):
        out.copy_(acc.to(out.dtype))
    elif mode == OUT_UNSHUFFLE16:
        out.copy_(_nhwc(F.pixel_unshuffle(acc.permute(0, 3, 1, 2), 2)).to(out.dtype))
    elif mode == OUT_SHUFFLE16:
        out.copy_(_nhwc(F.pixel_shuffle(acc.permute(0, 3, 1, 2), 2)).to(out.dtype))
    elif mode == OUT_FINAL_NCHW32:
        out.copy_(acc.permute(0, 3, 1, 2) + r["img"])
    else:
        raise AssertionError(mode)


def emu_dwconv(r):
    x, w, out = r["x"], r["w"], r["out"]
    cin = x.shape[-1]
    cw = w.float().t().reshape(cin, 1, 3, 3)
    y = F.conv2d(_nchw(x), cw, r["bias"], padding=1, groups=cin)
    if int(r["gate"]) == 2:                            # gate backward on the recomputed (fp32) stencil output
        c = cin // 2
        y1, y2, dg = y[:, :c], y[:, c:], _nchw(r["dg"])
        cdf = 0.5 * (1.0 + torch.erf(y1 * 0.7071067811865476))
        pdf = torch.exp(-0.5 * y1 * y1) * 0.3989422804014327
        y = torch.cat([dg * y2 * (cdf + y1 * pdf), dg * y1 * cdf], 1)
    elif r["gate"]:
        c = cin // 2
        y = F.gelu(y[:, :c]) * y[:, c:]
    out.copy_(_nhwc(y).to(out.dtype))


def emu_pwdw(r):
    """pir_pwdw: LayerNorm of x (no affine; rounded to the 16-bit type) -> 1x1 conv with the gamma-folded weights + t
    -> fp16 intermediate -> depthwise 3x3 (fp16 taps) -> optional GELU gate."""
    a, w, out = r["a"], r["w"], r["out"]
    B, H, W, K = a.shape
    npre = w.shape[0]
    A = a.float()
    if r["ln_mode"]:
        mu = A.mean(-1, keepdim=True)
        var = ((A * A).mean(-1, keepdim=True) - mu * mu).clamp_min(0)
        rstd = torch.rsqrt(var + 1e-5)
        A = A * rstd if r["ln_mode"] == LN_BIASFREE else (A - mu) * rstd
        A = A.to(a.dtype).float()
    pre = torch.einsum("bhwk,nk->bhwn", A, w.float()[:, :K])
    if r["vec_t"] is not None:
        pre = pre + r["vec_t"].view(1, 1, 1, -1)
    if a.dtype != torch.float32:                       # fp32 programs are the exact-wiring check
        pre = pre.clamp(-65504, 65504).to(torch.float16)
    out2 = r.get("out2")
    if out2 is None:
        emu_dwconv(dict(x=pre, w=r["dw_w"], out=out, bias=r["dw_bias"], gate=r["gate"]))
    else:                                              # two output tensors: channels [0, split) | [split, N)
        full = torch.empty(B, H, W, npre, dtype=out.dtype, device=out.device)
        emu_dwconv(dict(x=pre, w=r["dw_w"], out=full, bias=r["dw_bias"], gate=r["gate"]))
        out.copy_(full[..., :out.shape[-1]])
        out2.copy_(full[..., out.shape[-1]:])


def emu_mdta_finalize(r):
    heads, wo, wfold, temp = r["heads"], r["wo"], r["wfold"], r["temperature"]
    qkv = r["qk"] if r.get("qk") is not None else r["qkv"]
    B, H, W, cn = qkv.shape
    C = cn // 2 if r.get("qk") is not None else cn // 3
    c = C // heads
    q = qkv[..., :C].float().reshape(B, H * W, heads, c).permute(0, 2, 3, 1)          # [B, h, c, HW]
    k = qkv[..., C:2 * C].float().reshape(B, H * W, heads, c).permute(0, 2, 3, 1)
    g = q @ k.transpose(-1, -2)
    qn = q.norm(dim=-1).clamp_min(1e-12)
    kn = k.norm(dim=-1).clamp_min(1e-12)
    attn = torch.softmax(g / (qn[..., :, None] * kn[..., None, :]) * temp.view(1, heads, 1, 1), dim=-1)   # [B,h,c,c]
    wo_h = wo.view(C, heads, c)                                                        # [o, h, i]
    fold = torch.einsum("ohi,bhij->bohj", wo_h, attn).reshape(B, C, C)
    wfold.view(B, C, -1)[:, :, :C] = fold.to(wfold.dtype)


def emu_prompt(r):
    x, prm, lw, lb, out = r["x"], r["prompt"], r["lin_w"], r["lin_b"], r["out"]
    B, H, W, _ = x.shape
    emb = x.float().mean(dim=(1, 2))
    wts = torch.softmax(emb @ lw.t() + lb, dim=1)
    mix = torch.einsum("bl,lstd->bdst", wts, prm)
    out.copy_(_nhwc(F.interpolate(mix, (H, W), mode="bilinear", align_corners=bool(r.get("align_corners", False)))).to(out.dtype))


def emu_patch_embed(r):
    r["out"].copy_(_nhwc(F.conv2d(r["img"], r["w"], r["bias"], padding=1)).to(r["out"].dtype))


DISPATCH = {"gemm": emu_gemm, "dwconv": emu_dwconv, "pwdw": emu_pwdw, "mdta_gram": lambda r: None, "mdta_finalize": emu_mdta_finalize,
            "prompt": emu_prompt, "patch_embed": emu_patch_embed}


@torch.no_grad()
def run_program(engine, img):
    engine.img_in.copy_(img)
    for r in engine.ops:
        DISPATCH[r["kind"]](r)
    return engine.out.clone()


# ----------------------------------------------------------------------------------------------------
# training ops (promptir_b200/train_engine.py): forward extras and the backward kernels
# ----------------------------------------------------------------------------------------------------
def emu_ln_fwd(r):
    x = r["x"].float()
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(((x - mu) ** 2).mean(-1, keepdim=True) + 1e-5)
    xh = x * rstd if r["ln_mode"] == LN_BIASFREE else (x - mu) * rstd
    r["xhat"].copy_(xh.to(r["xhat"].dtype))
    r["rstd"].copy_(rstd.reshape(-1))


def emu_ln_bwd(r):
    d, xh, g = r["d"].float(), r["xhat"].float(), r["g"]
    rstd = r["rstd"].view(*d.shape[:3], 1)
    m = (d * xh).mean(-1, keepdim=True)
    if r["ln_mode"] == LN_BIASFREE:
        dx = rstd * (d - (xh - xh.mean(-1, keepdim=True)) * m)
    else:
        dx = rstd * (d - d.mean(-1, keepdim=True) - xh * m)
    g.copy_((g.float() + dx).to(g.dtype))


def _shifted(b, taps):
    """b[p + off(tap)] for every tap (zero outside the image).  tap = ky*3 + kx, off = (ky-1, kx-1)."""
    B, H, W, _ = b.shape
    if taps == 1:
        return [b]
    bp = F.pad(b, (0, 0, 1, 1, 1, 1))
    return [bp[:, t // 3:t // 3 + H, t % 3:t % 3 + W] for t in range(9)]


def _wg_views(ws, r):
    P, taps, M, N = r["P"], r["taps"], r["M"], r["N"]
    part = ws[:P * taps * M * N].view(P, taps, M, N)
    cs = ws[P * taps * M * N:P * taps * M * N + P * M].view(P, M) if r["colsum"] else None
    return part, cs


def emu_wgrad(r):
    a, b = r["a"].float(), r["b"].float()
    part, cs = _wg_views(r["ws"], r)
    part.zero_()
    sp = r["splits"]
    for t, bs in enumerate(_shifted(b, r["taps"])):
        if r["per_image"]:
            part[::sp, t] = torch.einsum("bhwm,bhwn->bmn", a, bs)
        else:
            part[0, t] = torch.einsum("bhwm,bhwn->mn", a, bs)
    if cs is not None:
        cs.zero_()
        if r["per_image"]:
            cs[::sp] = a.sum((1, 2))
        else:
            cs[0] = a.sum((0, 1, 2))


def _prow(R, half, half_pad, device):
    idx = torch.arange(R, device=device)
    return torch.where(idx < half, idx, idx - half + half_pad)


def emu_wgrad_fin(r):
    part, cs = _wg_views(r["ws"], r["wg"])
    dst = r["dst_w"]
    R, Cc = dst.shape[0], dst.shape[1]
    inv = r["inv_scale"]
    prow = _prow(R, r["half"], r["half_pad"], dst.device)
    G = part.sum(0)[:, prow, :Cc] * inv                     # [taps, R, Cc]
    s = cs.sum(0)[prow] * inv if cs is not None else None
    if r["gamma"] is not None:
        g0, w = G[0], r["w"].reshape(R, Cc).float()
        dW = r["gamma"].view(1, -1) * g0
        if r["beta"] is not None:
            dW = dW + r["beta"].view(1, -1) * s.view(-1, 1)
            r["dst_beta"].copy_((w * s.view(-1, 1)).sum(0))
        r["dst_gamma"].copy_((w * g0).sum(0))
        dst.copy_(dW.reshape(dst.shape))
    else:
        dst.copy_(G.permute(1, 2, 0).reshape(dst.shape))
    if r["dst_bias"] is not None:
        r["dst_bias"].copy_(s)


def emu_dw_wgrad(r):
    x, dy = r["x"].float(), r["dy"].float()
    dst = r["dst_w"]
    R = dst.shape[0]
    prow = _prow(R, r["half"], r["half_pad"], dst.device)
    taps = torch.stack([(xs * dy).sum((0, 1, 2)) for xs in _shifted(x, 9)])        # [9, Cp]
    dst.copy_((taps[:, prow].t() * r["inv_scale"]).reshape(dst.shape))
    if r["dst_bias"] is not None:
        r["dst_bias"].copy_(dy.sum((0, 1, 2))[prow] * r["inv_scale"])


def emu_gate_bwd(r):
    y, dg = r["y"], r["dgt"].float()
    c = dg.shape[-1]
    y1, y2 = y[..., :c].float(), y[..., c:].float()
    cdf = 0.5 * (1.0 + torch.erf(y1 * 0.7071067811865476))
    pdf = torch.exp(-0.5 * y1 * y1) * 0.3989422804014327
    d1 = dg * y2 * (cdf + y1 * pdf)
    d2 = dg * y1 * cdf
    y.copy_(torch.cat([d1, d2], -1).to(y.dtype))


def emu_mdta_bwd(r):
    qkv, heads, wo = r["qkv"], r["heads"], r["wo"]
    B, H, W, c3 = qkv.shape
    C = c3 // 3
    c = C // heads
    inv = r["inv_scale"]
    part, cs = _wg_views(r["ws"], r["wg"])
    dWf = part.view(B, -1, C, C).sum(1)                                               # [B, o, j']
    q = qkv[..., :C].float().reshape(B, H * W, heads, c).permute(0, 2, 3, 1)
    k = qkv[..., C:2 * C].float().reshape(B, H * W, heads, c).permute(0, 2, 3, 1)
    qn = q.norm(dim=-1).clamp_min(1e-12)
    kn = k.norm(dim=-1).clamp_min(1e-12)
    nn_ = qn[..., :, None] * kn[..., None, :]
    cos = (q @ k.transpose(-1, -2)) / nn_
    T = r["temperature"].float().view(1, heads, 1, 1)
    A = torch.softmax(cos * T, dim=-1)
    wo_h = wo.view(C, heads, c)
    dWf_h = dWf.view(B, C, heads, c)
    dA = torch.einsum("ohi,bohj->bhij", wo_h, dWf_h)
    r["dst_wo"].copy_(torch.einsum("bohj,bhij->ohi", dWf_h, A).reshape(C, C) * inv)
    dS = A * (dA - (dA * A).sum(-1, keepdim=True))
    r["dst_temp"].copy_((dS * cos).sum((0, 2, 3)) * inv)
    dcos = dS * T
    Mx = dcos / nn_
    rq = (dcos * cos).sum(-1) / (qn * qn)                                             # [B, h, i]
    rk = (dcos * cos).sum(-2) / (kn * kn)                                             # [B, h, j]
    full = torch.zeros(B, 2 * C, 2 * C, device=qkv.device)
    for h in range(heads):
        s = slice(h * c, (h + 1) * c)
        sk = slice(C + h * c, C + (h + 1) * c)
        full[:, s, sk] = Mx[:, h]
        full[:, sk, s] = Mx[:, h].transpose(-1, -2)
        full[:, s, s] = torch.diag_embed(-rq[:, h])
        full[:, sk, sk] = torch.diag_embed(-rk[:, h])
    wqk, wft = r["wqk"], r["wft"]
    wqk.view(B, 2 * C, -1)[:, :, :2 * C] = full.to(wqk.dtype)
    fold = torch.einsum("ohi,bhij->bohj", wo_h, A).reshape(B, C, C)
    wft.view(B, C, -1)[:, :, :C] = fold.transpose(1, 2).to(wft.dtype)
    if r["dst_bias"] is not None:
        r["dst_bias"].copy_(cs.sum(0) * inv)


def emu_shuffle(r):
    x, out = r["x"], r["out"]
    f = F.pixel_shuffle if r["up"] else F.pixel_unshuffle
    out.copy_(_nhwc(f(x.float().permute(0, 3, 1, 2), 2)).to(out.dtype))


def emu_prompt_train(r):
    """prompt forward that also leaves the pooled sums and softmax weights where pir_prompt_gen leaves them."""
    emu_prompt(r)
    x, lw, lb = r["x"], r["lin_w"], r["lin_b"]
    B, H, W, C = x.shape
    ws = r["ws"][:].view(B, -1, C)
    ws.zero_()
    ws[:, 0] = x.float().sum(dim=(1, 2))
    if r.get("weights_out") is not None:
        r["weights_out"].copy_(torch.softmax(x.float().mean(dim=(1, 2)) @ lw.t() + lb, dim=1))


def emu_prompt_bwd(r):
    dup, prm, wts, lw = r["dup"].float(), r["prompt"], r["weights"], r["lin_w"]
    B, H, W, D = dup.shape
    L, S = prm.shape[0], prm.shape[1]
    inv = r["inv_scale"]
    emb = r["pool_ws"].view(B, -1, r["C"]).sum(1) / r["HW"]
    with torch.enable_grad():
        mix = torch.zeros(B, D, S, S, device=dup.device, requires_grad=True)
        up = F.interpolate(mix, (H, W), mode="bilinear", align_corners=bool(r.get("align_corners", False)))
        (dmix,) = torch.autograd.grad(up, mix, dup.permute(0, 3, 1, 2))
    r["dst_prompt"].copy_(torch.einsum("bl,bdst->ldst", wts, dmix).unsqueeze(0) * inv)
    dw = torch.einsum("bdst,lstd->bl", dmix, prm)
    dlog = wts * (dw - (wts * dw).sum(1, keepdim=True))
    r["dst_lin_w"].copy_(dlog.t() @ emb * inv)
    r["dst_lin_b"].copy_(dlog.sum(0) * inv)
    r["demb"].copy_(dlog @ lw / r["HW"])


def emu_bcast_add(r):
    g = r["g"]
    g.copy_((g.float() + r["v"][:, None, None, :]).to(g.dtype))


def emu_to_nhwc16(r):
    src, out = r["src"], r["out"]
    out.zero_()
    out[..., :src.shape[1]] = (src.permute(0, 2, 3, 1) * r["scale"]).to(out.dtype)


DISPATCH.update({"ln_fwd": emu_ln_fwd, "ln_bwd": emu_ln_bwd, "wgrad": emu_wgrad, "wgrad_fin": emu_wgrad_fin, "dw_wgrad": emu_dw_wgrad,
                 "gate_bwd": emu_gate_bwd, "mdta_bwd": emu_mdta_bwd, "shuffle": emu_shuffle, "prompt_bwd": emu_prompt_bwd,
                 "bcast_add": emu_bcast_add, "to_nhwc16": emu_to_nhwc16})


@torch.no_grad()
def run_train(engine, img, d_out):
    """Emulated training forward + backward.  -> (out, {name: grad})."""
    engine.img_in.copy_(img)
    for r in engine.fwd_ops:
        (emu_prompt_train if r["kind"] == "prompt" else DISPATCH[r["kind"]])(r)
    out = engine.out.clone()
    engine.d_out.copy_(d_out)
    for r in engine.bwd_ops:
        DISPATCH[r["kind"]](r)
    return out, {k: v.clone() for k, v in engine.grads.items()}


def emu_ocab(r):
    """pir_ocab: the attention core of prompt_xrestormer.py:215-232 on the NHWC qkv tensor."""
    from oracle.xrestormer_oracle import ocab_core
    qkv, out = r["qkv"], r["out"]
    q, k, v = _nchw(qkv).chunk(3, dim=1)
    o = ocab_core(q.contiguous(), k.contiguous(), v.contiguous(), r["rel_h"], r["rel_w"], r["heads"])
    out.copy_(_nhwc(o).to(out.dtype))


DISPATCH["ocab"] = emu_ocab


def emu_ocab_bwd(r):
    """pir_ocab_bwd == autograd of the attention core (oracle.xrestormer_oracle.ocab_core) w.r.t. q, k, v and the two tables."""
    from oracle.xrestormer_oracle import ocab_core
    qkv, dout, dqkv = r["qkv"], r["dout"], r["dqkv"]
    with torch.enable_grad():
        x = _nchw(qkv).detach().clone().requires_grad_(True)
        rh, rw = r["rel_h"].detach().clone().requires_grad_(True), r["rel_w"].detach().clone().requires_grad_(True)
        q, k, v = x.chunk(3, dim=1)
        o = ocab_core(q, k, v, rh, rw, r["heads"])
        gx, grh, grw = torch.autograd.grad(o, (x, rh, rw), _nchw(dout))
    dqkv.copy_(_nhwc(gx).to(dqkv.dtype))
    r["dst_rel_h"].copy_(grh * r["inv_scale"])
    r["dst_rel_w"].copy_(grw * r["inv_scale"])


DISPATCH["ocab_bwd"] = emu_ocab_bwd

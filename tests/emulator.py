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
    if r["gate"]:
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
    emu_dwconv(dict(x=pre, w=r["dw_w"], out=out, bias=r["dw_bias"], gate=r["gate"]))


def emu_mdta_finalize(r):
    qkv, heads, wo, wfold, temp = r["qkv"], r["heads"], r["wo"], r["wfold"], r["temperature"]
    B, H, W, c3 = qkv.shape
    C = c3 // 3
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
    out.copy_(_nhwc(F.interpolate(mix, (H, W), mode="bilinear")).to(out.dtype))


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

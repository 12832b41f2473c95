"""GPU parity tests, kernel by kernel, through the C ABI (ctypes -> libpromptir_b200.so).

The expected value of every kernel is its specification restated in torch (tests/emulator.py) evaluated on the
same device buffers; the whole-network oracle checks live in test_gpu_model.py.  Run with `pytest -m gpu`.
"""
from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

import emulator
from promptir_b200 import ops, packing
from promptir_b200._lib import (LN_BIASFREE, LN_NONE, LN_WITHBIAS, OUT_FINAL_NCHW32, OUT_NHWC16, OUT_NHWC32,
                                OUT_SHUFFLE16, OUT_UNSHUFFLE16)

pytestmark = pytest.mark.gpu
DEV = "cuda"
DTYPES = [torch.float16, torch.bfloat16]


def stream():
    return torch.cuda.current_stream().cuda_stream


def tol(dt, scale=1.0):
    """(atol, rtol): one to two units in the last place of the 16-bit output plus fp32 summation-order noise."""
    return (2e-3 * scale, 2 ** -7) if dt == torch.bfloat16 else (3e-4 * scale, 2 ** -10)


def report_mismatch(name, got, ref, atol, rtol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    lim = atol + rtol * ref.abs()
    bad = err > lim
    if not bad.any():
        return
    idx = bad.nonzero()
    msg = [f"{name}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g} (ref absmax {ref.abs().max().item():.4g})"]
    for i in idx[:8].tolist():
        msg.append(f"  at {i}: got {got[tuple(i)].item():.6g} ref {ref[tuple(i)].item():.6g}")
    if got.dim() == 4:     # coarse error map over (pixel-tile, channel-block) to expose layout bugs
        B, H, W, Cc = got.shape
        e2 = err.reshape(B, H * W, Cc)
        rows = e2.amax(dim=2)[0]
        cols = e2.amax(dim=1)[0]
        msg.append("  row-block(16px) max err: " + " ".join(f"{v:.2g}" for v in rows[: 16 * 24].reshape(-1, 16).amax(1).tolist()))
        msg.append("  col-block(8ch) max err: " + " ".join(f"{v:.2g}" for v in F.pad(cols, (0, (-Cc) % 8)).reshape(-1, 8).amax(1).tolist()[:48]))
        msg.append(f"  nan got {int(torch.isnan(got).sum())} zero got {int((got == 0).sum())} zero ref {int((ref == 0).sum())}")
    pytest.fail("\n".join(msg))


def rand_act(B, H, W, C, dt, pitch=None, off=0, scale=1.0):
    pitch = pitch or C
    buf = (torch.randn(B, H, W, pitch, device=DEV) * scale).to(dt)
    return buf[..., off:off + C], buf


# --------------------------------------------------------------------------------------------------
# pointwise GEMM
# --------------------------------------------------------------------------------------------------
GEMM_CASES = [
    # B, H, W, K, N, a_pitch, a_off, ln, res, bias
    (1, 8, 16, 64, 64, None, 0, LN_NONE, False, False),       # one full tile, one k-block
    (1, 16, 16, 128, 128, None, 0, LN_NONE, False, False),    # two tiles, two k-blocks
    (2, 8, 24, 48, 144, None, 0, LN_WITHBIAS, False, False),  # K < 64, N split in two n-tiles, LN fold
    (1, 10, 10, 96, 96, None, 0, LN_NONE, True, False),       # partial M tile (100 px), residual
    (2, 16, 16, 96, 512, 192, 96, LN_WITHBIAS, False, True),  # A is a channel slice, 4 n-tiles, bias
    (1, 8, 8, 704, 192, None, 0, LN_NONE, False, False),      # 11 k-blocks -> ring wraps
    (1, 8, 8, 320, 960, None, 0, LN_BIASFREE, False, False),  # block_n 256 path (K >= 256)
    (1, 4, 4, 1024, 384, None, 0, LN_NONE, True, True),       # tiny image, deep K
    (3, 8, 8, 160, 96, 160, 0, LN_NONE, False, False),
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: "B%dH%dW%dK%dN%d_ln%d_r%d" % (c[0], c[1], c[2], c[3], c[4], c[7], c[8]))
def test_gemm_pointwise(case, dt):
    B, H, W, K, N, pitch, off, ln, use_res, use_bias = case
    torch.manual_seed(K * 131 + N)
    a, _ = rand_act(B, H, W, K, dt, pitch, off)
    if ln:
        a.add_(0.7)      # non-zero mean so the fold's mean term matters
    wt = torch.randn(N, K, device=DEV) / K ** 0.5
    gamma = torch.rand(K, device=DEV) + 0.5 if ln else None
    beta = torch.randn(K, device=DEV) * 0.1 if ln == LN_WITHBIAS else None
    bias = torch.randn(N, device=DEV) * 0.1 if use_bias else None
    w16, ln_s, vec_t = packing.pack_pointwise(wt, dt, gamma=gamma, beta=beta, bias=bias)
    out_buf = torch.zeros(B, H, W, N + 16, device=DEV, dtype=dt)
    out = out_buf[..., 8:8 + N]
    res = None
    if use_res:
        out.copy_(torch.randn(B, H, W, N, device=DEV).to(dt))
        res = out                                          # in place, like the residual stream
    rec = dict(a=a, w=w16, out=out, n=N, taps=1, out_mode=OUT_NHWC16, res=res, ln_mode=ln, ln_s=ln_s, vec_t=vec_t, img=None,
               w_batched=False)
    ref_out = out.clone()
    emulator.emu_gemm({**rec, "out": ref_out, "res": None if res is None else out.clone()})
    ops.gemm(a, w16, out, n=N, res=res, ln_mode=ln, ln_s=ln_s, vec_t=vec_t)(stream())
    torch.cuda.synchronize()
    atol, rtol = tol(dt, 2.0 if ln else 1.0)
    report_mismatch("gemm", out, ref_out, atol, rtol)
    assert float(out_buf[..., :8].abs().max()) == 0 and float(out_buf[..., 8 + N:].abs().max()) == 0, "wrote outside the slice"


@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_batched_weights(dt):
    """K4: per-image folded attention weights, A = v slice of the qkv tensor."""
    B, H, W, C = 3, 12, 12, 96
    torch.manual_seed(5)
    qkv, _ = rand_act(B, H, W, 3 * C, dt)
    x, _ = rand_act(B, H, W, C, dt)
    kp = packing.kpad_of(C)
    wf = torch.zeros(B, C, kp, device=DEV, dtype=dt)
    wf[:, :, :C] = (torch.randn(B, C, C, device=DEV) / C ** 0.5).to(dt)
    ref = x.clone()
    emulator.emu_gemm(dict(a=qkv[..., 2 * C:], w=wf, out=ref, n=C, taps=1, out_mode=OUT_NHWC16, res=x.clone(), ln_mode=0, ln_s=None,
                           vec_t=None, img=None, w_batched=True))
    ops.gemm(qkv[..., 2 * C:], wf, x, n=C, res=x, w_batched=True)(stream())
    torch.cuda.synchronize()
    report_mismatch("gemm_batched", x, ref, *tol(dt))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(6, 74, 128, 48), (7, 37, 128, 96), (20, 16, 24, 96), (5, 40, 40, 192)], ids=lambda c: "B%dH%dW%dC%d" % c)
def test_gemm_batched_weights_many_images_per_cta(case, dt):
    """More tiles than SMs: every persistent CTA walks three or more images, each with its own weight matrix."""
    B, H, W, C = case
    torch.manual_seed(C + B)
    qkv, _ = rand_act(B, H, W, 3 * C, dt)
    x, _ = rand_act(B, H, W, C, dt)
    kp = packing.kpad_of(C)
    wf = torch.zeros(B, C, kp, device=DEV, dtype=dt)
    wf[:, :, :C] = (torch.randn(B, C, C, device=DEV) / C ** 0.5).to(dt)
    ref = x.clone()
    emulator.emu_gemm(dict(a=qkv[..., 2 * C:], w=wf, out=ref, n=C, taps=1, out_mode=OUT_NHWC16, res=x.clone(), ln_mode=0, ln_s=None,
                           vec_t=None, img=None, w_batched=True))
    ops.gemm(qkv[..., 2 * C:], wf, x, n=C, res=x, w_batched=True)(stream())
    torch.cuda.synchronize()
    report_mismatch("gemm_batched", x, ref, *tol(dt))


def test_gemm_fp32_out_exact_small_integers():
    """Small-integer operands make every product and partial sum exact: the fp32 result must be bit-exact,
    which pins the UMMA/TMA descriptor layouts independently of any rounding."""
    for dt in DTYPES:
        B, H, W, K, N = 1, 16, 16, 192, 80
        torch.manual_seed(1)
        a = torch.randint(-4, 5, (B, H, W, K), device=DEV).to(dt)
        wt = torch.randint(-3, 4, (N, K), device=DEV).float()
        w16, _, _ = packing.pack_pointwise(wt, dt)
        out = torch.zeros(B, H, W, N, device=DEV)
        ops.gemm(a, w16, out, n=N, out_mode=OUT_NHWC32)(stream())
        torch.cuda.synchronize()
        ref = torch.einsum("bhwk,nk->bhwn", a.float(), wt)
        report_mismatch("gemm_exact", out, ref, 0.0, 0.0)


# --------------------------------------------------------------------------------------------------
# dense 3x3 (tap-shifted TMA) with folded PixelShuffle / PixelUnshuffle / final residual
# --------------------------------------------------------------------------------------------------
CONV_CASES = [
    # B, H, W, Cin, Cout, mode
    (1, 16, 16, 64, 64, OUT_NHWC16),
    (2, 8, 40, 48, 24, OUT_UNSHUFFLE16),
    (1, 12, 20, 96, 48, OUT_UNSHUFFLE16),
    (1, 6, 4, 192, 384, OUT_SHUFFLE16),
    (2, 16, 8, 96, 192, OUT_SHUFFLE16),
    (1, 24, 136, 96, 3, OUT_FINAL_NCHW32),
    (1, 4, 4, 320, 320, OUT_NHWC16),
    (1, 9, 5, 128, 128, OUT_NHWC16),
    # halo-tile kernel (conv3x3.cu, Cin <= 128): many items per CTA (buffer / phase wrap), ragged tiles, channel slices, 8-channel Cin
    (3, 64, 96, 48, 24, OUT_UNSHUFFLE16),
    (2, 50, 38, 96, 192, OUT_SHUFFLE16),
    (16, 32, 32, 64, 64, OUT_NHWC16),
    (2, 30, 22, 128, 128, OUT_NHWC16),
    (2, 128, 128, 96, 3, OUT_FINAL_NCHW32),
    (1, 20, 28, 8, 48, OUT_NHWC16),
    (1, 16, 16, 3 * 8, 16, OUT_NHWC32),
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "B%dH%dW%dC%dN%d_m%d" % c)
def test_gemm_conv3x3(case, dt):
    B, H, W, Cin, Cout, mode = case
    torch.manual_seed(Cin + Cout)
    a, _ = rand_act(B, H, W, Cin, dt, Cin + 8, 8)
    wt = torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cin) ** 0.5
    w16 = packing.pack_conv3x3(wt, dt)
    img = None
    if mode == OUT_NHWC16:
        out = torch.zeros(B, H, W, Cout, device=DEV, dtype=dt)
    elif mode == OUT_UNSHUFFLE16:
        out = torch.zeros(B, H // 2, W // 2, 4 * Cout + 8, device=DEV, dtype=dt)[..., 8:]
    elif mode == OUT_SHUFFLE16:
        out = torch.zeros(B, 2 * H, 2 * W, Cout // 4 + 8, device=DEV, dtype=dt)[..., :Cout // 4]
    else:
        out = torch.zeros(B, Cout, H, W, device=DEV)
        img = torch.rand(B, Cout, H, W, device=DEV)
    ref = out.clone()
    emulator.emu_gemm(dict(a=a, w=w16, out=ref, n=Cout, taps=9, out_mode=mode, res=None, ln_mode=0, ln_s=None, vec_t=None, img=img,
                           w_batched=False))
    ops.gemm(a, w16, out, n=Cout, taps=9, out_mode=mode, img=img)(stream())
    torch.cuda.synchronize()
    g, r = (out, ref) if mode != OUT_FINAL_NCHW32 else (out.permute(0, 2, 3, 1), ref.permute(0, 2, 3, 1))
    atol, rtol = tol(dt) if mode != OUT_FINAL_NCHW32 else (1e-5, 1e-5)
    report_mismatch("conv3x3", g, r, atol, rtol)


# --------------------------------------------------------------------------------------------------
# depthwise stencils
# --------------------------------------------------------------------------------------------------
DW_CASES = [(1, 8, 32, 64), (2, 16, 40, 144), (1, 13, 70, 288), (1, 4, 4, 576), (1, 32, 32, 480), (2, 9, 16, 2112)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", DW_CASES, ids=lambda c: "B%dH%dW%dC%d" % c)
@pytest.mark.parametrize("bias", [False, True])
def test_dwconv_plain(case, dt, bias):
    B, H, W, Cc = case
    torch.manual_seed(Cc)
    x, _ = rand_act(B, H, W, Cc, dt)
    wt = torch.randn(Cc, 1, 3, 3, device=DEV) / 3
    bv = torch.randn(Cc, device=DEV) * 0.1 if bias else None
    w16 = packing.pack_depthwise(wt, dt)
    out = torch.zeros(B, H, W, Cc, device=DEV, dtype=dt)
    ref = out.clone()
    emulator.emu_dwconv(dict(x=x, w=w16, out=ref, gate=False, bias=bv))
    ops.dwconv3x3(x, w16, out, gate=False, bias=bv)(stream())
    torch.cuda.synchronize()
    report_mismatch("dwconv_plain", out, ref, *tol(dt))


GATE_CASES = [(1, 8, 32, 32), (2, 16, 40, 128), (1, 13, 70, 256), (1, 4, 4, 1024), (1, 24, 24, 432), (1, 8, 8, 856), (1, 8, 8, 1872)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", GATE_CASES, ids=lambda c: "B%dH%dW%dC%d" % c)
def test_dwconv_gate(case, dt):
    B, H, W, hp = case
    torch.manual_seed(hp)
    x, _ = rand_act(B, H, W, 2 * hp, dt)
    wt = torch.randn(2 * hp, 1, 3, 3, device=DEV) / 3
    bv = torch.randn(2 * hp, device=DEV) * 0.1 if hp == 128 else None
    w16 = packing.pack_depthwise(wt, dt)
    out_buf = torch.zeros(B, H, W, hp + 8, device=DEV, dtype=dt)
    out = out_buf[..., :hp]
    ref = out.clone()
    emulator.emu_dwconv(dict(x=x, w=w16, out=ref, gate=True, bias=bv))
    ops.dwconv3x3(x, w16, out, gate=True, bias=bv)(stream())
    torch.cuda.synchronize()
    report_mismatch("dwconv_gate", out, ref, *tol(dt, 2.0))
    assert float(out_buf[..., hp:].abs().max()) == 0


# --------------------------------------------------------------------------------------------------
# fused LN -> 1x1 -> depthwise 3x3 (-> gate)
# --------------------------------------------------------------------------------------------------
PWDW_CASES = [
    # B, H, W, C, N, gate, ln, a_pitch, a_off, conv bias, dw bias
    (2, 40, 72, 48, 144, False, LN_WITHBIAS, None, 0, False, False),     # 8x32 tiles, ragged right/bottom edges, partial chunk
    (1, 16, 16, 96, 288, False, LN_WITHBIAS, None, 0, False, False),     # 16x16 tiles
    (2, 24, 40, 96, 256, True, LN_WITHBIAS, 160, 64, False, False),      # gate, x is a channel slice
    (1, 32, 32, 192, 512, True, LN_WITHBIAS, None, 0, False, False),     # 12x16 tiles (two M groups), three k-blocks
    (1, 20, 12, 192, 576, False, LN_BIASFREE, None, 0, True, True),      # BiasFree LN, both biases
    (1, 24, 24, 160, 432, True, LN_WITHBIAS, None, 0, True, True),       # hidden not a multiple of 32, K not a multiple of 64
    (1, 16, 24, 160, 480, False, LN_NONE, None, 0, False, False),        # no LayerNorm
    (3, 8, 8, 48, 128, True, LN_WITHBIAS, None, 0, False, False),        # tiny images, several per CTA
    (1, 64, 64, 96, 256, True, LN_WITHBIAS, None, 0, False, False),      # many tiles per CTA: ring/phase wrap
    # channel-major kernel (pwdwt.cu: gate, C <= 128, hidden % 128 == 0): ragged right/bottom tiles, biases, BiasFree, odd batch
    (2, 40, 72, 48, 128, True, LN_WITHBIAS, None, 0, True, True),
    (3, 20, 28, 96, 256, True, LN_BIASFREE, None, 0, False, True),
    (1, 8, 8, 96, 256, True, LN_NONE, None, 0, True, False),
    (5, 16, 16, 48, 128, True, LN_WITHBIAS, 96, 48, False, False),
    (2, 128, 128, 96, 256, True, LN_WITHBIAS, None, 0, False, False),
    # ... and its plain form (qkv): 144 = 128 + 16 channels, 288 = 256 + 32 (last block has one valid warp quarter)
    (2, 40, 72, 96, 288, False, LN_WITHBIAS, None, 0, True, True),
    (3, 20, 28, 48, 144, False, LN_BIASFREE, None, 0, False, True),
    (2, 128, 128, 96, 288, False, LN_WITHBIAS, None, 0, False, False),
    (1, 8, 8, 48, 144, False, LN_NONE, None, 0, True, False),
    # plain form, other channel splits: 192 = 128 + 64 (a half block, no replicated unit), 96 <= 128 (one block), 160 = 128 + 32 at
    # C = 48 (replicated unit of a full 32), 272 = 256 + 16 at C = 64, image height a multiple of neither 12 nor 4
    (2, 30, 40, 64, 192, False, LN_WITHBIAS, None, 0, True, True),
    (1, 24, 16, 32, 96, False, LN_BIASFREE, None, 0, False, False),
    (2, 26, 20, 48, 160, False, LN_WITHBIAS, None, 0, False, True),
    (1, 18, 50, 64, 272, False, LN_WITHBIAS, None, 0, True, False),
    (1, 22, 34, 64, 128, True, LN_WITHBIAS, None, 0, True, True),
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", PWDW_CASES, ids=lambda c: "B%dH%dW%dC%dN%d%s" % (c[0], c[1], c[2], c[3], c[4], "g" if c[5] else "p"))
def test_pwdw(case, dt):
    B, H, W, Cc, N, gate, ln, pitch, off, cbias, dbias = case
    torch.manual_seed(Cc + N)
    npre = 2 * N if gate else N
    assert ops.pwdw_supported(Cc, N, gate)
    x, _ = rand_act(B, H, W, Cc, dt, pitch, off, scale=2.0)
    x += 0.5                                                   # non-zero mean so the LayerNorm fold matters
    wt = torch.randn(npre, Cc, device=DEV) / Cc ** 0.5
    gamma = torch.rand(Cc, device=DEV) + 0.5 if ln else None
    beta = torch.randn(Cc, device=DEV) * 0.2 if ln == LN_WITHBIAS else None
    w16, ln_s, vec_t = packing.pack_pointwise(wt, dt, gamma=gamma, beta=beta, bias=torch.randn(npre, device=DEV) * 0.1 if cbias else None)
    dw = packing.pack_depthwise(torch.randn(npre, 1, 3, 3, device=DEV) / 3, torch.float16)
    db = torch.randn(npre, device=DEV) * 0.1 if dbias else None
    out_buf = torch.zeros(B, H, W, N + 8, device=DEV, dtype=dt)
    out = out_buf[..., :N]
    ref = out.clone()
    rec = dict(a=x, w=w16, dw_w=dw, out=ref, gate=gate, ln_mode=ln, vec_t=vec_t, dw_bias=db)
    emulator.emu_pwdw(rec)
    ops.pwdw(x, w16, dw, out, gate=gate, ln_mode=ln, vec_t=vec_t, dw_bias=db)(stream())
    torch.cuda.synchronize()
    # the nine taps are accumulated in fp16: up to ~3 roundings of 2^-11 relative to the largest partial sum (stated
    # tolerance of the fused kernel), on top of one unit in the last place of the 16-bit output
    atol, rtol = tol(dt, 2.0)
    report_mismatch("pwdw", out, ref, atol + 1.5e-3 * ref.abs().max().item(), rtol)
    assert float(out_buf[..., N:].abs().max()) == 0


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(2, 40, 56, 96), (3, 24, 24, 48), (1, 128, 128, 96)], ids=lambda c: "B%dH%dW%dC%d" % c)
def test_pwdw_split_outputs(case, dt):
    """qkv with q|k and v written to two dense tensors == the single-tensor result (model.py:121 chunk)."""
    B, H, W, Cc = case
    N = 3 * Cc
    assert ops.pwdw_split_supported(Cc, N)
    torch.manual_seed(Cc)
    x, _ = rand_act(B, H, W, Cc, dt, None, 0, scale=2.0)
    wt = torch.randn(N, Cc, device=DEV) / Cc ** 0.5
    w16, _, vec_t = packing.pack_pointwise(wt, dt, gamma=torch.rand(Cc, device=DEV) + 0.5, beta=torch.randn(Cc, device=DEV) * 0.2)
    dw = packing.pack_depthwise(torch.randn(N, 1, 3, 3, device=DEV) / 3, torch.float16)
    one = torch.zeros(B, H, W, N, device=DEV, dtype=dt)
    qk = torch.zeros(B, H, W, 2 * Cc, device=DEV, dtype=dt)
    v = torch.zeros(B, H, W, Cc, device=DEV, dtype=dt)
    ops.pwdw(x, w16, dw, one, gate=False, ln_mode=LN_WITHBIAS, vec_t=vec_t)(stream())
    ops.pwdw(x, w16, dw, qk, gate=False, ln_mode=LN_WITHBIAS, vec_t=vec_t, out2=v)(stream())
    torch.cuda.synchronize()
    assert torch.equal(one[..., :2 * Cc], qk) and torch.equal(one[..., 2 * Cc:], v)
    assert float(one.abs().max()) > 0


# --------------------------------------------------------------------------------------------------
# MDTA gram + finalize
# --------------------------------------------------------------------------------------------------
MDTA_CASES = [(2, 16, 16, 48, 1), (1, 32, 32, 96, 2), (2, 24, 24, 96, 1), (1, 16, 16, 192, 4), (2, 8, 8, 384, 8), (1, 8, 8, 704, 4),
              (1, 16, 16, 320, 4), (1, 24, 40, 160, 4), (1, 128, 128, 48, 1), (16, 32, 32, 192, 4), (3, 16, 16, 384, 8),
              (1, 64, 64, 384, 2), (1, 256, 256, 48, 1)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", MDTA_CASES, ids=lambda c: "B%dH%dW%dC%dh%d" % c)
def test_mdta(case, dt):
    B, H, W, Cc, heads = case
    torch.manual_seed(Cc + heads)
    qkv, _ = rand_act(B, H, W, 3 * Cc, dt)
    temp = torch.rand(heads, device=DEV) * 4 + 0.5
    wo = (torch.randn(Cc, Cc, device=DEV) / Cc ** 0.5).contiguous()
    kp = packing.kpad_of(Cc)
    wfold = torch.zeros(B, Cc, kp, device=DEV, dtype=dt)
    splits = ops.mdta_splits(B, H * W, Cc)
    ws = torch.zeros(ops.mdta_ws_floats(B, Cc, splits), device=DEV)
    gram, fin = ops.mdta(qkv, heads, ws, temp, wo, wfold, splits)
    gram(stream())
    fin(stream())
    torch.cuda.synchronize()
    ref = torch.zeros_like(wfold)
    emulator.emu_mdta_finalize(dict(qkv=qkv, heads=heads, wo=wo, wfold=ref, temperature=temp))
    # check the raw Gram partial sums too (summed over splits) for a sharper diagnosis.  Workspace layout:
    # [B][splits][C][c] diagonal head blocks, then [B][splits][2][C] squared norms (csrc/mdta.cu)
    c = Cc // heads
    g = ws[: B * splits * Cc * c].view(B, splits, Cc, c).sum(1)
    q = qkv[..., :Cc].float().reshape(B, -1, Cc)
    k = qkv[..., Cc:2 * Cc].float().reshape(B, -1, Cc)
    gfull = torch.einsum("bpi,bpj->bij", q, k)
    gref = torch.stack([gfull[:, h * c:(h + 1) * c, h * c:(h + 1) * c] for h in range(heads)], 1).reshape(B, Cc, c)
    gerr = (g - gref).abs().max().item()
    assert gerr <= 1e-3 * max(1.0, gref.abs().max().item()), f"gram mismatch {gerr} (ref max {gref.abs().max().item()})"
    nrm = ws[B * splits * Cc * c: B * splits * (Cc * c + 2 * Cc)].view(B, splits, 2, Cc).sum(1)
    nref = torch.stack([q.pow(2).sum(1), k.pow(2).sum(1)], dim=1)
    assert torch.allclose(nrm, nref, rtol=1e-4, atol=1e-3), f"norm mismatch {(nrm - nref).abs().max().item()}"
    report_mismatch("mdta_wfold", wfold.view(B, 1, Cc, kp), ref.view(B, 1, Cc, kp), *tol(dt))
    # the attention matrices kept in the workspace for pir_mdta_bwd: [B][heads][c][c] after the partials
    attn = ws[B * splits * (Cc * c + 2 * Cc):][: B * Cc * c].view(B, heads, c, c)
    qn = torch.nn.functional.normalize(q.transpose(1, 2).reshape(B, heads, c, -1), dim=-1)
    kn = torch.nn.functional.normalize(k.transpose(1, 2).reshape(B, heads, c, -1), dim=-1)
    aref = ((qn @ kn.transpose(-1, -2)) * temp.view(1, heads, 1, 1)).softmax(-1)
    assert (attn - aref).abs().max().item() <= 2e-5, f"attention mismatch {(attn - aref).abs().max().item()}"


_FIN_AB = """
import sys, torch
sys.path.insert(0, %r)
from promptir_b200 import ops, packing
torch.manual_seed(5)
out = []
for (B, H, W, Cc, heads) in [(2, 16, 16, 48, 1), (1, 24, 40, 160, 4), (2, 8, 8, 704, 4), (3, 16, 16, 384, 8)]:
    qkv = torch.randn(B, H, W, 3 * Cc, device="cuda").to(torch.bfloat16)
    temp = torch.rand(heads, device="cuda") * 4 + 0.5
    wo = (torch.randn(Cc, Cc, device="cuda") / Cc ** 0.5).contiguous()
    wfold = torch.zeros(B, Cc, packing.kpad_of(Cc), device="cuda", dtype=torch.bfloat16)
    splits = ops.mdta_splits(B, H * W, Cc)
    ws = torch.zeros(ops.mdta_ws_floats(B, Cc, splits), device="cuda")
    gram, fin = ops.mdta(qkv, heads, ws, temp, wo, wfold, splits)
    s = torch.cuda.current_stream().cuda_stream
    gram(s); fin(s)
    torch.cuda.synchronize()
    out += [wfold.float().cpu(), ws.cpu(), torch.tensor([float(fin.kernels)])]
torch.save(out, sys.argv[1])
"""


def test_mdta_finalize_fused_equals_two_kernel_path(tmp_path):
    """PIR_MDTA_FUSED=0 selects the softmax + fold kernels; the fused kernel keeps their arithmetic and summation order,
    so folded weights and saved attention must agree bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for flag in ("1", "0"):
        f = str(tmp_path / f"fin{flag}.pt")
        env = dict(os.environ, PIR_MDTA_FUSED=flag)
        subprocess.run([sys.executable, "-c", _FIN_AB % root, f], check=True, env=env, timeout=300)
        res.append(torch.load(f))
    assert [int(t.item()) for t in res[0][2::3]] == [1, 1, 1, 1] and [int(t.item()) for t in res[1][2::3]] == [2, 2, 2, 2]
    for i, (a, b) in enumerate(zip(res[0], res[1])):
        if i % 3 != 2:
            assert torch.equal(a, b), f"case {i // 3}: {'wfold' if i % 3 == 0 else 'workspace'} differs"


# --------------------------------------------------------------------------------------------------
# prompt generation, patch embed, tile blend
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(2, 16, 16, 384, 320, 16), (1, 64, 64, 192, 128, 32), (2, 8, 24, 96, 64, 64), (1, 40, 24, 96, 64, 64),
                                  (16, 32, 32, 384, 320, 16), (16, 128, 128, 96, 64, 64), (3, 136, 72, 96, 64, 64), (16, 64, 64, 192, 128, 32)],
                         ids=lambda c: "B%dH%dW%dC%dD%dS%d" % c)
def test_prompt_gen(case, dt):
    """Single-launch PromptGenBlock (device-wide barrier between pool and mix): upsampling x2, identity, downsampling, ragged tiles,
    more work items than CTAs; launched three times on the same sync counters (CUDA-graph replays re-arm the barrier)."""
    B, H, W, Cc, D, S = case
    torch.manual_seed(D)
    x, _ = rand_act(B, H, W, Cc, dt, Cc + D, 0)
    prm = packing.pack_prompt(torch.rand(1, 5, D, S, S, device=DEV))
    lw = (torch.randn(5, Cc, device=DEV) / Cc ** 0.5).contiguous()
    lb = torch.randn(5, device=DEV) * 0.1
    out = torch.zeros(B, H, W, D, device=DEV, dtype=dt)
    ws = torch.zeros(ops.prompt_ws_floats(B, H * W, Cc), device=DEV)
    wts = torch.zeros(B, 5, device=DEV)
    launch = ops.prompt_gen(x, prm, lw, lb, out, ws, wts)
    assert launch.kernels == 1
    for _ in range(3):
        out.zero_()
        launch(stream())
    torch.cuda.synchronize()
    ref = torch.zeros_like(out)
    emulator.emu_prompt(dict(x=x, prompt=prm, lin_w=lw, lin_b=lb, out=ref))
    wref = torch.softmax(x.float().mean(dim=(1, 2)) @ lw.t() + lb, dim=1)
    assert torch.allclose(wts, wref, atol=1e-5), (wts, wref)
    report_mismatch("prompt", out, ref, *tol(dt))


@pytest.mark.xfail(strict=False, reason="written after round 2's GPU budget was spent: not run on a B200 yet.  An XPASS in the log is the "
                                       "first measurement; a failure would point at the align_corners=True up-sampling branch of pir_prompt_gen")
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [(1, 256, 256, 96, 64, 64), (1, 64, 64, 384, 320, 16)], ids=lambda c: "B%dH%dW%dC%dD%dS%d" % c)
def test_prompt_gen_align_corners_upsampling(case, dt):
    """PromptXRestormer's PromptBlock resizes with align_corners=True (prompt_xrestormer.py:340), and at BASELINE.json configs[4]
    (512x512) it up-samples (x4, x4, x4).  The GPU goldens of that network (64, 64x128, 128) only down-scale or keep the size, and
    test_prompt_gen covers up-sampling for align_corners=False only: this is the missing combination, on exactly the prompt1 / prompt3
    launch shapes of that configuration (which bench.py's xrestormer sub-record has run, checked on a 128x128 crop only)."""
    B, H, W, Cc, D, S = case
    torch.manual_seed(D + 1)
    x, _ = rand_act(B, H, W, Cc, dt, Cc + D, 0)
    prm = packing.pack_prompt(torch.rand(1, 5, D, S, S, device=DEV))
    lw = (torch.randn(5, Cc, device=DEV) / Cc ** 0.5).contiguous()
    lb = torch.randn(5, device=DEV) * 0.1
    out = torch.zeros(B, H, W, D, device=DEV, dtype=dt)
    ws = torch.zeros(ops.prompt_ws_floats(B, H * W, Cc), device=DEV)
    ops.prompt_gen(x, prm, lw, lb, out, ws, None, align_corners=True)(stream())
    torch.cuda.synchronize()
    ref = torch.zeros_like(out)
    emulator.emu_prompt(dict(x=x, prompt=prm, lin_w=lw, lin_b=lb, out=ref, align_corners=True))
    report_mismatch("prompt(align_corners=True, up)", out, ref, *tol(dt))


@pytest.mark.parametrize("dt", DTYPES)
def test_patch_embed(dt):
    B, H, W = 2, 24, 40
    torch.manual_seed(0)
    img = torch.rand(B, 3, H, W, device=DEV)
    wt = (torch.randn(48, 3, 3, 3, device=DEV) / 27 ** 0.5).contiguous()
    buf = torch.zeros(B, H, W, 96, device=DEV, dtype=dt)
    out = buf[..., 48:]
    ops.patch_embed(img, wt, None, out)(stream())
    torch.cuda.synchronize()
    ref = torch.zeros_like(out)
    emulator.emu_patch_embed(dict(img=img, w=wt, bias=None, out=ref))
    report_mismatch("patch_embed", out, ref, *tol(dt))
    assert float(buf[..., :48].abs().max()) == 0


def test_tile_blend_matches_reference_loop():
    from oracle.promptir_oracle import tile_origins, tiled_restore
    torch.manual_seed(0)
    Cc, H, W, tile, ov = 3, 72, 56, 32, 8
    img = torch.rand(1, Cc, H, W, device=DEV)
    fn = lambda t: t * 1.5 - 0.2 + 0.1 * t.flip(-1)
    ys, xs = tile_origins(H, tile, ov), tile_origins(W, tile, ov)
    tiles = torch.cat([fn(img[..., y:y + tile, x:x + tile]) for y in ys for x in xs]).contiguous()
    out = torch.zeros(Cc, H, W, device=DEV)
    ops.tile_blend(tiles, torch.tensor(ys, device=DEV, dtype=torch.int32), torch.tensor(xs, device=DEV, dtype=torch.int32), out, stream())
    torch.cuda.synchronize()
    ref = tiled_restore(fn, img, tile, ov)[0]
    assert torch.equal(out, ref) or (out - ref).abs().max().item() < 1e-6

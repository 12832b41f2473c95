"""Thin Python wrappers over the C ABI: they turn torch tensors (used only as device memory) into the plain
pointer/size descriptors of include/promptir_b200.h.  Each wrapper returns a *prepared launch* (callable
taking a raw stream handle) so an engine can build its program once and replay it with no per-call
marshalling.  Activations are passed as NHWC views `[B, H, W, C]` with unit channel stride; a view may be a
channel slice of a wider buffer (that is how concatenations are folded away).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_FP16, LN_BIASFREE, LN_NONE, LN_WITHBIAS, OUT_FINAL_NCHW32, OUT_NHWC16,
                   OUT_NHWC32, OUT_SHUFFLE16, OUT_UNSHUFFLE16)

Launch = Callable[[int], None]


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return DTYPE_BF16
    if dt == torch.float16:
        return DTYPE_FP16
    raise ValueError(f"promptir_b200 computes in bfloat16 or float16, got {dt}")


def _nhwc(t: torch.Tensor, what: str):
    """-> (ptr, B, H, W, C, pitch, bstride) of an NHWC view."""
    if t.dim() != 4 or t.stride(3) != 1 or t.stride(1) != t.shape[2] * t.stride(2):
        raise ValueError(f"{what}: expected an NHWC view with unit channel stride and dense rows, got "
                         f"shape {tuple(t.shape)} strides {t.stride()}")
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is not on a CUDA device (promptir_b200 has no CPU path)")
    return t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3], t.stride(2), t.stride(0)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _prepared(fn_name: str, desc, keep, kernels: int = 1) -> Launch:
    lib = _lib.load()
    fn = getattr(lib, fn_name)
    ref = C.byref(desc)

    def launch(stream: int, _fn=fn, _ref=ref, _keep=(desc, keep), _name=fn_name, _k=kernels) -> None:
        rc = _fn(_ref, stream)
        if rc:
            _lib.check(rc, _name)
        _lib.launch_count += _k                  # kernels this entry point enqueues

    launch.desc = desc
    launch.kernels = kernels
    return launch


# ----------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, n: int, taps: int = 1, out_mode: int = OUT_NHWC16,
         res: Optional[torch.Tensor] = None, ln_mode: int = LN_NONE, ln_s: Optional[torch.Tensor] = None,
         vec_t: Optional[torch.Tensor] = None, img: Optional[torch.Tensor] = None, w_batched: bool = False) -> Launch:
    """1x1 / 3x3 conv as tcgen05 GEMM (pir_gemm).  `w` is the packed K-major weight (see packing.py)."""
    pa, B, H, W, K, apitch, abs_ = _nhwc(a, "gemm.a")
    d = _lib.PirGemm()
    d.dtype = dtype_code(a.dtype)
    d.B, d.H, d.W, d.K, d.N, d.taps = B, H, W, K, n, taps
    d.w_batched = int(w_batched)
    d.out_mode, d.ln_mode = out_mode, ln_mode
    d.a, d.a_pitch, d.a_bstride = pa, apitch, abs_
    d.w = w.data_ptr()
    if out_mode == OUT_FINAL_NCHW32:
        assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, n, H, W)
        d.out, d.out_pitch, d.out_bstride = out.data_ptr(), 1, n * H * W
        assert img is not None and img.dtype == torch.float32 and img.is_contiguous() and img.shape == out.shape
    else:
        po, oB, oH, oW, oC, opitch, obs = _nhwc(out, "gemm.out")
        want = {OUT_NHWC16: (H, W, n), OUT_NHWC32: (H, W, n), OUT_UNSHUFFLE16: (H // 2, W // 2, 4 * n),
                OUT_SHUFFLE16: (2 * H, 2 * W, n // 4)}[out_mode]
        if (oH, oW, oC) != want or oB != B:
            raise ValueError(f"gemm.out: shape {tuple(out.shape)} does not match mode {out_mode} (want {want})")
        if (out_mode == OUT_NHWC32) != (out.dtype == torch.float32):
            raise ValueError("gemm.out: dtype does not match the store mode")
        d.out, d.out_pitch, d.out_bstride = po, opitch, obs
    if res is not None:
        pr, rB, rH, rW, rC, rpitch, rbs = _nhwc(res, "gemm.res")
        assert (rB, rH, rW, rC) == (B, H, W, n) and res.dtype == a.dtype
        d.res, d.res_pitch, d.res_bstride = pr, rpitch, rbs
    d.ln_s, d.vec_t, d.img = _ptr(ln_s), _ptr(vec_t), _ptr(img)
    return _prepared("pir_gemm", d, (a, w, out, res, ln_s, vec_t, img))


def dwconv3x3(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, gate, bias: Optional[torch.Tensor] = None,
              dg: Optional[torch.Tensor] = None) -> Launch:
    """gate: False plain, True GELU gate, 2 = gate backward with the stencil recomputed (needs dg; out has as many channels as x)."""
    px, B, H, W, Cin, xp, xbs = _nhwc(x, "dwconv.in")
    po, oB, oH, oW, Cout, op, obs = _nhwc(out, "dwconv.out")
    gate = int(gate)
    if gate == 2:
        assert Cin == Cout and Cin % 2 == 0 and dg is not None
        Cout = Cin // 2
    assert (oB, oH, oW) == (B, H, W) and Cin == (2 * Cout if gate else Cout)
    d = _lib.PirDwConv()
    d.dtype, d.gate = dtype_code(x.dtype), gate
    d.B, d.H, d.W, d.C = B, H, W, Cout
    if dg is not None:
        pg, gB, gH, gW, gC, gp, gbs = _nhwc(dg, "dwconv.dg")
        assert (gB, gH, gW, gC) == (B, H, W, Cout) and dg.dtype == x.dtype
        d.dg, d.dg_pitch, d.dg_bstride = pg, gp, gbs
    d.in_, d.in_pitch, d.in_bstride = px, xp, xbs
    d.w, d.bias = w.data_ptr(), _ptr(bias)
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    return _prepared("pir_dwconv3x3", d, (x, w, out, bias, dg))


def pwdw_supported(c: int, n: int, gate: bool) -> bool:
    return bool(_lib.load().pir_pwdw_supported(c, n, int(gate)))


def pwdw_split_supported(c: int, n: int) -> bool:
    return bool(_lib.load().pir_pwdw_split_supported(c, n))


def pwdw(x: torch.Tensor, w: torch.Tensor, dw_w: torch.Tensor, out: torch.Tensor, *, gate: bool, ln_mode: int = LN_NONE,
         vec_t: Optional[torch.Tensor] = None, dw_bias: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None) -> Launch:
    """Fused LN -> 1x1 conv -> depthwise 3x3 (-> GELU gate) (pir_pwdw).  w: packed [Npre, Kpad]; dw_w: fp16 [9, Npre].
    out2 (gate == False only): the channels after out's go to this second tensor (q|k and v of MDTA as two dense tensors)."""
    pa, B, H, W, K, apitch, abs_ = _nhwc(x, "pwdw.a")
    po, oB, oH, oW, N, opitch, obs = _nhwc(out, "pwdw.out")
    split = 0
    if out2 is not None:
        assert not gate
        po2, o2B, o2H, o2W, N2, o2pitch, o2bs = _nhwc(out2, "pwdw.out2")
        assert (o2B, o2H, o2W) == (B, H, W) and out2.dtype == x.dtype
        split, N = N, N + N2
    npre = 2 * N if gate else N
    assert (oB, oH, oW) == (B, H, W) and out.dtype == x.dtype
    assert dw_w.dtype == torch.float16 and dw_w.is_contiguous() and tuple(dw_w.shape) == (9, npre)
    assert w.dtype == x.dtype and w.is_contiguous() and w.shape[0] == npre and w.shape[1] == (K + 63) // 64 * 64
    d = _lib.PirPwDw()
    d.dtype, d.gate, d.ln_mode = dtype_code(x.dtype), int(gate), ln_mode
    d.B, d.H, d.W, d.C, d.N = B, H, W, K, N
    d.a, d.a_pitch, d.a_bstride = pa, apitch, abs_
    d.w, d.vec_t = w.data_ptr(), _ptr(vec_t)
    d.dw_w, d.dw_bias = dw_w.data_ptr(), _ptr(dw_bias)
    d.out, d.out_pitch, d.out_bstride = po, opitch, obs
    if out2 is not None:
        d.out2, d.out2_pitch, d.out2_bstride, d.split = po2, o2pitch, o2bs, split
    return _prepared("pir_pwdw", d, (x, w, dw_w, out, vec_t, dw_bias, out2))


def mdta_splits(B: int, HW: int, Cdim: int) -> int:
    return int(_lib.load().pir_mdta_splits(B, HW, Cdim))


def mdta_ws_floats(B: int, Cdim: int, splits: int) -> int:
    return int(_lib.load().pir_mdta_ws_floats(B, Cdim, splits))


def mdta(qkv: torch.Tensor, heads: int, ws: torch.Tensor, temperature: torch.Tensor, wo: torch.Tensor,
         wfold: torch.Tensor, splits: int, qk_only: bool = False):
    """-> (gram_launch, finalize_launch).  qkv: NHWC [B,H,W,3C] after the depthwise conv (only q|k = channels [0, 2C) are read), or,
    with qk_only, the dense q|k tensor [B,H,W,2C] of the two-tensor qkv layout."""
    pq, B, H, W, C3, qp, qbs = _nhwc(qkv, "mdta.qkv")
    Cdim = C3 // 2 if qk_only else C3 // 3
    assert ws.dtype == torch.float32 and ws.numel() >= mdta_ws_floats(B, Cdim, splits)
    assert wo.dtype == torch.float32 and wo.is_contiguous() and wo.numel() == Cdim * Cdim
    assert temperature.dtype == torch.float32 and temperature.numel() == heads
    kpad = (Cdim + 63) // 64 * 64
    assert wfold.is_contiguous() and wfold.numel() == B * Cdim * kpad and wfold.dtype == qkv.dtype
    d = _lib.PirMdta()
    d.dtype = dtype_code(qkv.dtype)
    d.B, d.HW, d.C, d.heads, d.splits = B, H * W, Cdim, heads, splits
    d.qkv, d.qkv_pitch, d.qkv_bstride = pq, qp, qbs
    d.ws, d.temperature, d.wo, d.wfold = ws.data_ptr(), temperature.data_ptr(), wo.data_ptr(), wfold.data_ptr()
    keep = (qkv, ws, temperature, wo, wfold)
    return (_prepared("pir_mdta_gram", d, keep),
            _prepared("pir_mdta_finalize", d, keep, kernels=int(_lib.load().pir_mdta_finalize_kernels(Cdim, heads))))


def prompt_ws_floats(B: int, HW: int, Cdim: int) -> int:
    return int(_lib.load().pir_prompt_ws_floats(B, HW, Cdim))


def prompt_gen(x: torch.Tensor, prompt: torch.Tensor, lin_w: torch.Tensor, lin_b: torch.Tensor, out: torch.Tensor,
               ws: torch.Tensor, weights_out: Optional[torch.Tensor] = None, align_corners: bool = False) -> Launch:
    """prompt: fp32 [L, S, S, D] (components, channels last)."""
    px, B, H, W, Cdim, xp, xbs = _nhwc(x, "prompt.x")
    po, oB, oH, oW, D, op, obs = _nhwc(out, "prompt.out")
    L, S, S2, D2 = prompt.shape
    assert (oB, oH, oW) == (B, H, W) and S == S2 and D == D2 and prompt.is_contiguous() and prompt.dtype == torch.float32
    assert lin_w.dtype == torch.float32 and lin_w.is_contiguous() and tuple(lin_w.shape) == (L, Cdim)
    assert ws.numel() >= prompt_ws_floats(B, H * W, Cdim)
    d = _lib.PirPrompt()
    d.dtype = dtype_code(x.dtype)
    d.B, d.H, d.W, d.C, d.L, d.D, d.S = B, H, W, Cdim, L, D, S
    d.x, d.x_pitch, d.x_bstride = px, xp, xbs
    d.prompt, d.lin_w, d.lin_b = prompt.data_ptr(), lin_w.data_ptr(), lin_b.data_ptr()
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    d.ws, d.weights_out = ws.data_ptr(), _ptr(weights_out)
    d.align_corners = int(align_corners)
    # two zero-initialised ints owned by this call site: the arrive counters of the single-launch kernel's device-wide barrier
    sync = torch.zeros(2, dtype=torch.int32, device=x.device)
    d.sync = sync.data_ptr()
    return _prepared("pir_prompt_gen", d, (x, prompt, lin_w, lin_b, out, ws, weights_out, sync),
                     kernels=int(_lib.load().pir_prompt_gen_kernels(C.byref(d))))


def patch_embed(img: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor) -> Launch:
    assert img.dtype == torch.float32 and img.is_contiguous() and img.is_cuda
    B, Cin, H, W = img.shape
    po, oB, oH, oW, Cout, op, obs = _nhwc(out, "patch_embed.out")
    assert (oB, oH, oW) == (B, H, W) and w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape) == (Cout, Cin, 3, 3)
    d = _lib.PirPatchEmbed()
    d.dtype = dtype_code(out.dtype)
    d.B, d.H, d.W, d.Cin, d.Cout = B, H, W, Cin, Cout
    d.img, d.w, d.bias = img.data_ptr(), w.data_ptr(), _ptr(bias)
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    return _prepared("pir_patch_embed", d, (img, w, bias, out))


def tile_blend(tiles: torch.Tensor, ys: torch.Tensor, xs: torch.Tensor, out: torch.Tensor, stream: int) -> None:
    """tiles fp32 [ny*nx, C, th, tw]; ys/xs int32 device vectors of origins; out fp32 [C, H, W]."""
    T, Cc, th, tw = tiles.shape
    ny, nx = ys.numel(), xs.numel()
    assert T == ny * nx and tiles.is_contiguous() and out.is_contiguous() and tiles.dtype == out.dtype == torch.float32
    assert ys.dtype == xs.dtype == torch.int32 and ys.is_cuda and xs.is_cuda and tiles.is_cuda and out.is_cuda
    lib = _lib.load()
    _lib.check(lib.pir_tile_blend(tiles.data_ptr(), ny, nx, ys.data_ptr(), xs.data_ptr(), Cc, th, tw, out.data_ptr(),
                                  out.shape[-2], out.shape[-1], stream), "pir_tile_blend")
    _lib.launch_count += 1


# ----------------------------------------------------------------------------------------------------
# training ops (see train_engine.py).  The multi-output ones take the engine's record dict.
# ----------------------------------------------------------------------------------------------------
def _f32ptr(t: Optional[torch.Tensor], what: str) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{what}: expected a contiguous float32 tensor, got {t.dtype} strides {t.stride()}")
    return t.data_ptr()


def _ln_desc(x, xhat, rstd, ln_mode, g=None):
    px, B, H, W, Cc, xp, xbs = _nhwc(x, "ln.x")
    ph, hB, hH, hW, hC, hp, hbs = _nhwc(xhat, "ln.xhat")
    assert (hB, hH, hW, hC) == (B, H, W, Cc) and xhat.dtype == x.dtype and rstd.numel() == B * H * W
    d = _lib.PirLn()
    d.dtype, d.ln_mode = dtype_code(x.dtype), ln_mode
    d.B, d.H, d.W, d.C = B, H, W, Cc
    d.x, d.x_pitch, d.x_bstride = px, xp, xbs
    d.xhat, d.xh_pitch, d.xh_bstride = ph, hp, hbs
    d.rstd = _f32ptr(rstd, "ln.rstd")
    if g is not None:
        pg, gB, gH, gW, gC, gp, gbs = _nhwc(g, "ln.g")
        assert (gB, gH, gW, gC) == (B, H, W, Cc) and g.dtype == x.dtype
        d.g, d.g_pitch, d.g_bstride = pg, gp, gbs
    return d


def ln_fwd(x, xhat, rstd, ln_mode) -> Launch:
    return _prepared("pir_ln_fwd", _ln_desc(x, xhat, rstd, ln_mode), (x, xhat, rstd))


def ln_bwd(d, xhat, rstd, g, ln_mode) -> Launch:
    """g += LayerNorm backward of d = dL/d(xhat)."""
    return _prepared("pir_ln_bwd", _ln_desc(d, xhat, rstd, ln_mode, g), (d, xhat, rstd, g))


def wgrad_splits(B: int, HW: int, M: int, N: int, taps: int, per_image: bool) -> int:
    return int(_lib.load().pir_wgrad_splits(B, HW, M, N, taps, int(per_image)))


def _wg_layout(rec):
    """float offsets of (partials, colsum) of a wgrad record inside the shared workspace."""
    n = rec["P"] * rec["taps"] * rec["M"] * rec["N"]
    return 0, (n if rec["colsum"] else None)


def wgrad(ws: torch.Tensor, rec: dict) -> Launch:
    a, b = rec["a"], rec["b"]
    pa, B, H, W, M, ap, abs_ = _nhwc(a, "wgrad.a")
    pb, bB, bH, bW, N, bp, bbs = _nhwc(b, "wgrad.b")
    assert (bB, bH, bW) == (B, H, W) and a.dtype == b.dtype and (M, N) == (rec["M"], rec["N"])
    off_p, off_c = _wg_layout(rec)
    need = rec["P"] * rec["taps"] * M * N + (rec["P"] * M if rec["colsum"] else 0)
    assert ws.dtype == torch.float32 and ws.numel() >= need
    d = _lib.PirWgrad()
    d.dtype = dtype_code(a.dtype)
    d.B, d.H, d.W, d.M, d.N, d.taps = B, H, W, M, N, rec["taps"]
    d.per_image, d.splits = int(rec["per_image"]), rec["splits"]
    d.a, d.a_pitch, d.a_bstride = pa, ap, abs_
    d.b, d.b_pitch, d.b_bstride = pb, bp, bbs
    d.ws = ws.data_ptr() + 4 * off_p
    d.colsum = None if off_c is None else ws.data_ptr() + 4 * off_c
    return _prepared("pir_wgrad", d, (a, b, ws))


def wgrad_finalize(ws: torch.Tensor, rec: dict) -> Launch:
    wg, dst = rec["wg"], rec["dst_w"]
    off_p, off_c = _wg_layout(wg)
    d = _lib.PirWgradFin()
    d.P, d.M, d.N, d.taps = wg["P"], wg["M"], wg["N"], wg["taps"]
    d.R, d.Cc, d.half, d.half_pad = dst.shape[0], dst.shape[1], rec["half"], rec["half_pad"]
    assert dst.numel() == d.R * d.Cc * d.taps
    d.ws = ws.data_ptr() + 4 * off_p
    d.colsum = None if off_c is None else ws.data_ptr() + 4 * off_c
    d.inv_scale = rec["inv_scale"]
    d.gamma, d.beta, d.w = _f32ptr(rec["gamma"], "fin.gamma"), _f32ptr(rec["beta"], "fin.beta"), _f32ptr(rec["w"], "fin.w")
    d.dst_w = _f32ptr(dst, "fin.dst_w")
    d.dst_gamma, d.dst_beta, d.dst_bias = (_f32ptr(rec[k], "fin." + k) for k in ("dst_gamma", "dst_beta", "dst_bias"))
    keep = (ws, dst, rec["gamma"], rec["beta"], rec["w"], rec["dst_gamma"], rec["dst_beta"], rec["dst_bias"])
    return _prepared("pir_wgrad_finalize", d, keep, kernels=2 if (rec["gamma"] is not None or rec["dst_bias"] is not None) else 1)


def dw_wgrad_parts(B: int, H: int, W: int, Cp: int) -> int:
    return int(_lib.load().pir_dw_wgrad_parts(B, H, W, Cp))


def dw_wgrad(ws: torch.Tensor, rec: dict) -> Launch:
    x, dy, dst = rec["x"], rec["dy"], rec["dst_w"]
    px, B, H, W, Cp, xp, xbs = _nhwc(x, "dw_wgrad.x")
    pd, dB, dH, dW, dC, dp, dbs = _nhwc(dy, "dw_wgrad.dy")
    assert (dB, dH, dW, dC) == (B, H, W, Cp) and x.dtype == dy.dtype and ws.numel() >= rec["parts"] * 10 * Cp
    d = _lib.PirDwWgrad()
    d.dtype = dtype_code(x.dtype)
    d.B, d.H, d.W, d.C = B, H, W, Cp
    d.parts, d.R, d.half, d.half_pad = rec["parts"], dst.shape[0], rec["half"], rec["half_pad"]
    assert dst.numel() == d.R * 9
    d.x, d.x_pitch, d.x_bstride = px, xp, xbs
    d.dy, d.dy_pitch, d.dy_bstride = pd, dp, dbs
    d.ws, d.inv_scale = ws.data_ptr(), rec["inv_scale"]
    d.dst_w, d.dst_bias = _f32ptr(dst, "dw_wgrad.dst_w"), _f32ptr(rec["dst_bias"], "dw_wgrad.dst_bias")
    return _prepared("pir_dw_wgrad", d, (x, dy, ws, dst, rec["dst_bias"]), kernels=2)


def gate_bwd(y: torch.Tensor, dgt: torch.Tensor) -> Launch:
    py, B, H, W, C2, yp, ybs = _nhwc(y, "gate_bwd.y")
    pg, gB, gH, gW, Cc, gp, gbs = _nhwc(dgt, "gate_bwd.dg")
    assert (gB, gH, gW) == (B, H, W) and C2 == 2 * Cc and y.dtype == dgt.dtype
    d = _lib.PirGateBwd()
    d.dtype = dtype_code(y.dtype)
    d.B, d.H, d.W, d.C = B, H, W, Cc
    d.y, d.y_pitch, d.y_bstride = py, yp, ybs
    d.dg, d.dg_pitch, d.dg_bstride = pg, gp, gbs
    return _prepared("pir_gate_bwd", d, (y, dgt))


def mdta_bwd_ws_floats(B: int, Cdim: int, heads: int) -> int:
    return int(_lib.load().pir_mdta_bwd_ws_floats(B, Cdim, heads))


def mdta_bwd(ws: torch.Tensor, rec: dict) -> Launch:
    wg = rec["wg"]
    B, Cdim, heads = rec["B"], rec["C"], rec["heads"]
    off_p, off_c = _wg_layout(wg)
    off_s = wg["P"] * Cdim * Cdim + (wg["P"] * Cdim if wg["colsum"] else 0)
    assert ws.numel() >= off_s + mdta_bwd_ws_floats(B, Cdim, heads)
    wft, wqk = rec["wft"], rec["wqk"]
    assert wft.is_contiguous() and wqk.is_contiguous() and wft.numel() == B * Cdim * ((Cdim + 63) // 64 * 64)
    assert wqk.numel() == B * 2 * Cdim * ((2 * Cdim + 63) // 64 * 64)
    d = _lib.PirMdtaBwd()
    d.dtype = dtype_code(wft.dtype)
    d.B, d.C, d.heads, d.splits_f, d.splits_b = B, Cdim, heads, rec["splits"], wg["splits"]
    d.ws_f = _f32ptr(rec["fws"], "mdta_bwd.fws")
    d.ws_b = ws.data_ptr() + 4 * off_p
    d.colsum_b = None if off_c is None else ws.data_ptr() + 4 * off_c
    d.temperature, d.wo = _f32ptr(rec["temperature"].reshape(-1), "mdta_bwd.temperature"), _f32ptr(rec["wo"], "mdta_bwd.wo")
    d.inv_scale = rec["inv_scale"]
    d.scratch = ws.data_ptr() + 4 * off_s
    d.wft, d.wqk = wft.data_ptr(), wqk.data_ptr()
    d.dst_wo, d.dst_temp, d.dst_bias = _f32ptr(rec["dst_wo"], "dst_wo"), _f32ptr(rec["dst_temp"], "dst_temp"), _f32ptr(rec["dst_bias"], "dst_bias")
    keep = (ws, rec["fws"], rec["temperature"], rec["wo"], wft, wqk, rec["dst_wo"], rec["dst_temp"], rec["dst_bias"])
    return _prepared("pir_mdta_bwd", d, keep, kernels=6)


def pixel_shuffle(x: torch.Tensor, out: torch.Tensor, *, up: bool) -> Launch:
    """up: x [B,H,W,4c] -> out [B,2H,2W,c];  not up: x [B,2H,2W,c] -> out [B,H,W,4c]."""
    px, xB, xH, xW, xC, xp, xbs = _nhwc(x, "shuffle.in")
    po, oB, oH, oW, oC, op, obs = _nhwc(out, "shuffle.out")
    big = (xH, xW, xC) if up else (oH, oW, oC)
    small = (oH, oW, oC) if up else (xH, xW, xC)
    assert xB == oB and small == (2 * big[0], 2 * big[1], big[2] // 4) and x.dtype == out.dtype
    d = _lib.PirShuffle()
    d.dtype, d.up = dtype_code(x.dtype), int(up)
    d.B, d.H, d.W, d.C = xB, big[0], big[1], big[2]
    d.in_, d.in_pitch, d.in_bstride = px, xp, xbs
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    return _prepared("pir_pixel_shuffle", d, (x, out))


def prompt_bwd_ws_floats(B: int, L: int, D: int, S: int) -> int:
    return int(_lib.load().pir_prompt_bwd_ws_floats(B, L, D, S))


def prompt_bwd(ws: torch.Tensor, rec: dict) -> Launch:
    dup, prm = rec["dup"], rec["prompt"]
    pd, B, H, W, D, dp, dbs = _nhwc(dup, "prompt_bwd.dup")
    L, S = prm.shape[0], prm.shape[1]
    assert tuple(prm.shape) == (L, S, S, D) and ws.numel() >= prompt_bwd_ws_floats(B, L, D, S)
    assert rec["dst_prompt"].numel() == prm.numel() and tuple(rec["demb"].shape) == (B, rec["C"])
    d = _lib.PirPromptBwd()
    d.dtype = dtype_code(dup.dtype)
    d.B, d.H, d.W, d.C, d.L, d.D, d.S = B, H, W, rec["C"], L, D, S
    d.dup, d.dup_pitch, d.dup_bstride = pd, dp, dbs
    d.prompt, d.weights = _f32ptr(prm, "prompt"), _f32ptr(rec["weights"], "weights")
    d.pool_ws, d.lin_w = _f32ptr(rec["pool_ws"], "pool_ws"), _f32ptr(rec["lin_w"], "lin_w")
    d.inv_scale = rec["inv_scale"]
    d.align_corners = int(bool(rec.get("align_corners", False)))
    d.scratch, d.demb = ws.data_ptr(), _f32ptr(rec["demb"], "demb")
    d.dst_prompt, d.dst_lin_w, d.dst_lin_b = (_f32ptr(rec[k], k) for k in ("dst_prompt", "dst_lin_w", "dst_lin_b"))
    keep = (ws, dup, prm, rec["weights"], rec["pool_ws"], rec["lin_w"], rec["demb"], rec["dst_prompt"], rec["dst_lin_w"], rec["dst_lin_b"])
    return _prepared("pir_prompt_bwd", d, keep, kernels=4)


def bcast_add(g: torch.Tensor, v: torch.Tensor) -> Launch:
    pg, B, H, W, Cc, gp, gbs = _nhwc(g, "bcast_add.g")
    assert tuple(v.shape) == (B, Cc)
    d = _lib.PirBcastAdd()
    d.dtype = dtype_code(g.dtype)
    d.B, d.H, d.W, d.C = B, H, W, Cc
    d.g, d.g_pitch, d.g_bstride = pg, gp, gbs
    d.v = _f32ptr(v, "bcast_add.v")
    return _prepared("pir_bcast_add", d, (g, v))


def nchw32_to_nhwc16(src: torch.Tensor, out: torch.Tensor, scale: float) -> Launch:
    assert src.dtype == torch.float32 and src.is_contiguous() and src.is_cuda
    B, Cc, H, W = src.shape
    po, oB, oH, oW, oC, op, obs = _nhwc(out, "to_nhwc16.out")
    assert (oB, oH, oW, oC) == (B, H, W, 8) and Cc <= 8
    d = _lib.PirToNhwc16()
    d.dtype = dtype_code(out.dtype)
    d.B, d.C, d.H, d.W, d.Cpad = B, Cc, H, W, 8
    d.src = src.data_ptr()
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    d.scale = scale
    return _prepared("pir_nchw32_to_nhwc16", d, (src, out))


def ocab(qkv: torch.Tensor, rel_h: torch.Tensor, rel_w: torch.Tensor, out: torch.Tensor, *, heads: int, dim_head: int = 16, ws: int = 8,
         ows: int = 12) -> Launch:
    """Overlapping cross-attention core (pir_ocab): qkv NHWC [B,H,W,3*heads*dim_head] -> out NHWC [B,H,W,heads*dim_head]."""
    pq, B, H, W, C3, qp, qbs = _nhwc(qkv, "ocab.qkv")
    po, oB, oH, oW, oC, op, obs = _nhwc(out, "ocab.out")
    inner = heads * dim_head
    assert C3 == 3 * inner and (oB, oH, oW, oC) == (B, H, W, inner) and out.dtype == qkv.dtype
    assert tuple(rel_h.shape) == tuple(rel_w.shape) == (2 * ows - 1, dim_head)
    d = _lib.PirOcab()
    d.dtype = dtype_code(qkv.dtype)
    d.B, d.H, d.W, d.heads, d.dim_head, d.ws, d.ows = B, H, W, heads, dim_head, ws, ows
    d.qkv, d.qkv_pitch, d.qkv_bstride = pq, qp, qbs
    d.rel_h, d.rel_w = _f32ptr(rel_h, "ocab.rel_h"), _f32ptr(rel_w, "ocab.rel_w")
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    return _prepared("pir_ocab", d, (qkv, rel_h, rel_w, out))


def ocab_bwd_ws_floats(B: int, H: int, W: int, heads: int) -> int:
    return int(_lib.load().pir_ocab_bwd_ws_floats(B, H, W, heads))


def ocab_bwd(ws: torch.Tensor, rec: dict) -> Launch:
    """OCAB backward (pir_ocab_bwd): rec holds qkv, dout, rel_h, rel_w, dqkv, dst_rel_h, dst_rel_w, heads, inv_scale."""
    qkv, dout, dqkv = rec["qkv"], rec["dout"], rec["dqkv"]
    pq, B, H, W, C3, qp, qbs = _nhwc(qkv, "ocab_bwd.qkv")
    pd, dB, dH, dW, inner, dp, dbs = _nhwc(dout, "ocab_bwd.dout")
    pg, gB, gH, gW, gC, gp, gbs = _nhwc(dqkv, "ocab_bwd.dqkv")
    heads = rec["heads"]
    assert inner == 16 * heads and C3 == 3 * inner and (dB, dH, dW) == (B, H, W) and (gB, gH, gW, gC) == (B, H, W, C3)
    assert ws.numel() >= ocab_bwd_ws_floats(B, H, W, heads)
    d = _lib.PirOcabBwd()
    d.dtype = dtype_code(qkv.dtype)
    d.B, d.H, d.W, d.heads, d.dim_head, d.ws_, d.ows = B, H, W, heads, 16, 8, 12
    d.qkv, d.qkv_pitch, d.qkv_bstride = pq, qp, qbs
    d.dout, d.dout_pitch, d.dout_bstride = pd, dp, dbs
    d.rel_h, d.rel_w = _f32ptr(rec["rel_h"], "rel_h"), _f32ptr(rec["rel_w"], "rel_w")
    d.dqkv, d.dqkv_pitch, d.dqkv_bstride = pg, gp, gbs
    d.ws, d.inv_scale = ws.data_ptr(), rec["inv_scale"]
    d.dst_rel_h, d.dst_rel_w = _f32ptr(rec["dst_rel_h"], "dst_rel_h"), _f32ptr(rec["dst_rel_w"], "dst_rel_w")
    return _prepared("pir_ocab_bwd", d, (ws, qkv, dout, dqkv, rec["rel_h"], rec["rel_w"], rec["dst_rel_h"], rec["dst_rel_w"]), kernels=3)

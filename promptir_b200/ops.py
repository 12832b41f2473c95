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


def dwconv3x3(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, gate: bool, bias: Optional[torch.Tensor] = None) -> Launch:
    px, B, H, W, Cin, xp, xbs = _nhwc(x, "dwconv.in")
    po, oB, oH, oW, Cout, op, obs = _nhwc(out, "dwconv.out")
    assert (oB, oH, oW) == (B, H, W) and Cin == (2 * Cout if gate else Cout)
    d = _lib.PirDwConv()
    d.dtype, d.gate = dtype_code(x.dtype), int(gate)
    d.B, d.H, d.W, d.C = B, H, W, Cout
    d.in_, d.in_pitch, d.in_bstride = px, xp, xbs
    d.w, d.bias = w.data_ptr(), _ptr(bias)
    d.out, d.out_pitch, d.out_bstride = po, op, obs
    return _prepared("pir_dwconv3x3", d, (x, w, out, bias))


def pwdw_supported(c: int, n: int, gate: bool) -> bool:
    return bool(_lib.load().pir_pwdw_supported(c, n, int(gate)))


def pwdw(x: torch.Tensor, w: torch.Tensor, dw_w: torch.Tensor, out: torch.Tensor, *, gate: bool, ln_mode: int = LN_NONE,
         vec_t: Optional[torch.Tensor] = None, dw_bias: Optional[torch.Tensor] = None) -> Launch:
    """Fused LN -> 1x1 conv -> depthwise 3x3 (-> GELU gate) (pir_pwdw).  w: packed [Npre, Kpad]; dw_w: fp16 [9, Npre]."""
    pa, B, H, W, K, apitch, abs_ = _nhwc(x, "pwdw.a")
    po, oB, oH, oW, N, opitch, obs = _nhwc(out, "pwdw.out")
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
    return _prepared("pir_pwdw", d, (x, w, dw_w, out, vec_t, dw_bias))


def mdta_splits(B: int, HW: int, Cdim: int) -> int:
    return int(_lib.load().pir_mdta_splits(B, HW, Cdim))


def mdta_ws_floats(B: int, Cdim: int, splits: int) -> int:
    return int(_lib.load().pir_mdta_ws_floats(B, Cdim, splits))


def mdta(qkv: torch.Tensor, heads: int, ws: torch.Tensor, temperature: torch.Tensor, wo: torch.Tensor,
         wfold: torch.Tensor, splits: int):
    """-> (gram_launch, finalize_launch).  qkv: NHWC [B,H,W,3C] after the depthwise conv."""
    pq, B, H, W, C3, qp, qbs = _nhwc(qkv, "mdta.qkv")
    Cdim = C3 // 3
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
    return _prepared("pir_mdta_gram", d, keep), _prepared("pir_mdta_finalize", d, keep, kernels=2)


def prompt_ws_floats(B: int, HW: int, Cdim: int) -> int:
    return int(_lib.load().pir_prompt_ws_floats(B, HW, Cdim))


def prompt_gen(x: torch.Tensor, prompt: torch.Tensor, lin_w: torch.Tensor, lin_b: torch.Tensor, out: torch.Tensor,
               ws: torch.Tensor, weights_out: Optional[torch.Tensor] = None) -> Launch:
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
    return _prepared("pir_prompt_gen", d, (x, prompt, lin_w, lin_b, out, ws, weights_out), kernels=2)


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

"""ctypes binding of libpromptir_b200.so (the C ABI declared in include/promptir_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised -- the product path never silently runs anything else.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PROMPTIR_B200_LIB: another build of the same library (same-box A/B timing of kernel changes); default = the in-tree build
LIB_PATH = os.environ.get("PROMPTIR_B200_LIB") or os.path.join(_HERE, "libpromptir_b200.so")

DTYPE_FP16, DTYPE_BF16 = 0, 1
OUT_NHWC16, OUT_UNSHUFFLE16, OUT_SHUFFLE16, OUT_FINAL_NCHW32, OUT_NHWC32 = 0, 1, 2, 3, 4
LN_NONE, LN_WITHBIAS, LN_BIASFREE = 0, 1, 2

i32, i64, vp = C.c_int32, C.c_int64, C.c_void_p


class PirGemm(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("K", i32), ("N", i32), ("taps", i32),
                ("w_batched", i32), ("out_mode", i32), ("ln_mode", i32),
                ("a", vp), ("a_pitch", i64), ("a_bstride", i64),
                ("w", vp),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64),
                ("res", vp), ("res_pitch", i64), ("res_bstride", i64),
                ("ln_s", vp), ("vec_t", vp), ("img", vp)]


class PirDwConv(C.Structure):
    _fields_ = [("dtype", i32), ("gate", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32),
                ("in_", vp), ("in_pitch", i64), ("in_bstride", i64),
                ("w", vp), ("bias", vp),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64),
                ("dg", vp), ("dg_pitch", i64), ("dg_bstride", i64)]


class PirPwDw(C.Structure):
    _fields_ = [("dtype", i32), ("gate", i32), ("ln_mode", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("N", i32),
                ("a", vp), ("a_pitch", i64), ("a_bstride", i64),
                ("w", vp), ("vec_t", vp), ("dw_w", vp), ("dw_bias", vp),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64),
                ("out2", vp), ("out2_pitch", i64), ("out2_bstride", i64), ("split", i32)]


class PirMdta(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("HW", i32), ("C", i32), ("heads", i32), ("splits", i32),
                ("qkv", vp), ("qkv_pitch", i64), ("qkv_bstride", i64),
                ("ws", vp), ("temperature", vp), ("wo", vp), ("wfold", vp)]


class PirPrompt(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("L", i32), ("D", i32), ("S", i32),
                ("x", vp), ("x_pitch", i64), ("x_bstride", i64),
                ("prompt", vp), ("lin_w", vp), ("lin_b", vp),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64),
                ("ws", vp), ("weights_out", vp), ("align_corners", i32), ("sync", vp)]


class PirOcab(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("heads", i32), ("dim_head", i32), ("ws", i32), ("ows", i32),
                ("qkv", vp), ("qkv_pitch", i64), ("qkv_bstride", i64),
                ("rel_h", vp), ("rel_w", vp),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64)]


class PirOcabBwd(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("heads", i32), ("dim_head", i32), ("ws_", i32), ("ows", i32),
                ("qkv", vp), ("qkv_pitch", i64), ("qkv_bstride", i64),
                ("dout", vp), ("dout_pitch", i64), ("dout_bstride", i64),
                ("rel_h", vp), ("rel_w", vp),
                ("dqkv", vp), ("dqkv_pitch", i64), ("dqkv_bstride", i64),
                ("ws", vp), ("inv_scale", C.c_float),
                ("dst_rel_h", vp), ("dst_rel_w", vp)]


class PirPatchEmbed(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32),
                ("img", vp), ("w", vp), ("bias", vp),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64)]


f32 = C.c_float


class PirLn(C.Structure):
    _fields_ = [("dtype", i32), ("ln_mode", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32),
                ("x", vp), ("x_pitch", i64), ("x_bstride", i64),
                ("xhat", vp), ("xh_pitch", i64), ("xh_bstride", i64),
                ("rstd", vp),
                ("g", vp), ("g_pitch", i64), ("g_bstride", i64)]


class PirWgrad(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("M", i32), ("N", i32), ("taps", i32),
                ("per_image", i32), ("splits", i32),
                ("a", vp), ("a_pitch", i64), ("a_bstride", i64),
                ("b", vp), ("b_pitch", i64), ("b_bstride", i64),
                ("ws", vp), ("colsum", vp)]


class PirWgradFin(C.Structure):
    _fields_ = [("P", i32), ("M", i32), ("N", i32), ("taps", i32), ("R", i32), ("Cc", i32), ("half", i32), ("half_pad", i32),
                ("ws", vp), ("colsum", vp), ("inv_scale", f32),
                ("gamma", vp), ("beta", vp), ("w", vp),
                ("dst_w", vp), ("dst_gamma", vp), ("dst_beta", vp), ("dst_bias", vp)]


class PirDwWgrad(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("parts", i32), ("R", i32), ("half", i32),
                ("half_pad", i32),
                ("x", vp), ("x_pitch", i64), ("x_bstride", i64),
                ("dy", vp), ("dy_pitch", i64), ("dy_bstride", i64),
                ("ws", vp), ("inv_scale", f32), ("dst_w", vp), ("dst_bias", vp)]


class PirGateBwd(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32),
                ("y", vp), ("y_pitch", i64), ("y_bstride", i64),
                ("dg", vp), ("dg_pitch", i64), ("dg_bstride", i64)]


class PirMdtaBwd(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("C", i32), ("heads", i32), ("splits_f", i32), ("splits_b", i32),
                ("ws_f", vp), ("ws_b", vp), ("colsum_b", vp), ("temperature", vp), ("wo", vp),
                ("inv_scale", f32), ("scratch", vp), ("wft", vp), ("wqk", vp),
                ("dst_wo", vp), ("dst_temp", vp), ("dst_bias", vp)]


class PirShuffle(C.Structure):
    _fields_ = [("dtype", i32), ("up", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32),
                ("in_", vp), ("in_pitch", i64), ("in_bstride", i64),
                ("out", vp), ("out_pitch", i64), ("out_bstride", i64)]


class PirPromptBwd(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("L", i32), ("D", i32), ("S", i32),
                ("dup", vp), ("dup_pitch", i64), ("dup_bstride", i64),
                ("prompt", vp), ("weights", vp), ("pool_ws", vp), ("lin_w", vp),
                ("inv_scale", f32), ("scratch", vp), ("demb", vp),
                ("dst_prompt", vp), ("dst_lin_w", vp), ("dst_lin_b", vp), ("align_corners", i32)]


class PirBcastAdd(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("H", i32), ("W", i32), ("C", i32),
                ("g", vp), ("g_pitch", i64), ("g_bstride", i64), ("v", vp)]


class PirToNhwc16(C.Structure):
    _fields_ = [("dtype", i32), ("B", i32), ("C", i32), ("H", i32), ("W", i32), ("Cpad", i32),
                ("src", vp), ("out", vp), ("out_pitch", i64), ("out_bstride", i64), ("scale", f32)]


class PirPackJob(C.Structure):
    _fields_ = [("kind", i32), ("dst_dtype", i32), ("n", i32), ("k", i32), ("transpose", i32), ("flip", i32),
                ("n_total", i32), ("k_pad", i32), ("row_split", i32), ("row_hp", i32), ("col_split", i32), ("col_hp", i32),
                ("gamma_axis", i32), ("reserved", i32),
                ("src", vp), ("gamma", vp), ("beta", vp), ("bias", vp),
                ("dst", vp), ("ln_s", vp), ("vec_t", vp)]


PACK_POINTWISE, PACK_CONV3X3, PACK_DEPTHWISE, PACK_VEC, PACK_PROMPT = 0, 1, 2, 3, 4

# every symbol include/promptir_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pir_abi_version": (i32, []),
    "pir_last_error": (C.c_char_p, []),
    "pir_check_device": (i32, []),
    "pir_gemm": (i32, [C.POINTER(PirGemm), vp]),
    "pir_dwconv3x3": (i32, [C.POINTER(PirDwConv), vp]),
    "pir_pwdw_supported": (i32, [i32, i32, i32]),
    "pir_pwdw_split_supported": (i32, [i32, i32]),
    "pir_pwdw": (i32, [C.POINTER(PirPwDw), vp]),
    "pir_mdta_splits": (i32, [i32, i32, i32]),
    "pir_mdta_ws_floats": (i64, [i32, i32, i32]),
    "pir_mdta_finalize_kernels": (i32, [i32, i32]),
    "pir_mdta_gram": (i32, [C.POINTER(PirMdta), vp]),
    "pir_mdta_finalize": (i32, [C.POINTER(PirMdta), vp]),
    "pir_prompt_ws_floats": (i64, [i32, i32, i32]),
    "pir_prompt_gen": (i32, [C.POINTER(PirPrompt), vp]),
    "pir_prompt_gen_kernels": (i32, [C.POINTER(PirPrompt)]),
    "pir_patch_embed": (i32, [C.POINTER(PirPatchEmbed), vp]),
    "pir_tile_blend": (i32, [vp, i32, i32, vp, vp, i32, i32, i32, vp, i32, i32, vp]),
    "pir_ln_fwd": (i32, [C.POINTER(PirLn), vp]),
    "pir_ln_bwd": (i32, [C.POINTER(PirLn), vp]),
    "pir_wgrad_splits": (i32, [i32, i32, i32, i32, i32, i32]),
    "pir_wgrad": (i32, [C.POINTER(PirWgrad), vp]),
    "pir_wgrad_finalize": (i32, [C.POINTER(PirWgradFin), vp]),
    "pir_dw_wgrad_parts": (i32, [i32, i32, i32, i32]),
    "pir_dw_wgrad": (i32, [C.POINTER(PirDwWgrad), vp]),
    "pir_gate_bwd": (i32, [C.POINTER(PirGateBwd), vp]),
    "pir_mdta_bwd_ws_floats": (i64, [i32, i32, i32]),
    "pir_mdta_bwd": (i32, [C.POINTER(PirMdtaBwd), vp]),
    "pir_pixel_shuffle": (i32, [C.POINTER(PirShuffle), vp]),
    "pir_prompt_bwd_ws_floats": (i64, [i32, i32, i32, i32]),
    "pir_prompt_bwd": (i32, [C.POINTER(PirPromptBwd), vp]),
    "pir_bcast_add": (i32, [C.POINTER(PirBcastAdd), vp]),
    "pir_nchw32_to_nhwc16": (i32, [C.POINTER(PirToNhwc16), vp]),
    "pir_ocab": (i32, [C.POINTER(PirOcab), vp]),
    "pir_ocab_bwd_ws_floats": (i64, [i32, i32, i32, i32]),
    "pir_ocab_bwd": (i32, [C.POINTER(PirOcabBwd), vp]),
    "pir_mirror_pad": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "pir_psnr_ssim_ws_bytes": (i64, [i32, i32, i32, i32]),
    "pir_psnr_ssim": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    "pir_add_noise": (i32, [vp, vp, i64, f32, C.c_uint64, vp]),
    "pir_repack_rows": (i64, [C.POINTER(PirPackJob)]),
    "pir_repack": (i32, [vp, vp, i32, i32, vp]),
}

_lib = None
_lock = threading.Lock()
launch_count = 0          # kernels enqueued through this binding (bench.py reports it as gpu_launches)


def load() -> C.CDLL:
    """Load the library once; raise loudly if it is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"promptir_b200: native library not found at {LIB_PATH}. Build it with "
                    "`make -C promptir_b200/csrc` (needs nvcc, sm_100a). There is no fallback path.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)           # AttributeError if the symbol is missing -> loud
                fn.restype = res
                fn.argtypes = args
            if lib.pir_abi_version() != 1:
                raise RuntimeError("promptir_b200: ABI version mismatch between _lib.py and the shared library")
            _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().pir_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"promptir_b200.{what} failed ({code}): {msg}")

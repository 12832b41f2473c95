"""Training program of PromptIR on one B200: forward that keeps what the backward needs, and the hand-derived backward.

Reference: `PromptIRModel.training_step` (train.py:37-46) runs `net(x)` under autograd and calls `loss.backward()`; autograd
then replays ~5000 ATen backward kernels.  Here both directions are fixed launch programs over the C ABI
(include/promptir_b200.h), built once per (batch, height, width, dtype) like the inference `Engine`:

forward (per TransformerBlock, net/model.py:192-196) -- the unfused kernels, because the backward needs the pre-stencil tensors:
    ln_fwd   x -> xhat (16-bit, no affine: gamma is folded into the 1x1 weights, beta into its additive vector), rstd
    gemm     xhat . Wg + t -> qkv_pre                       dwconv   -> qkv
    mdta_gram / mdta_finalize -> Wfold[b]                   gemm     v . Wfold[b] + x -> x
    ln_fwd   x -> xhat2, rstd2                              gemm     xhat2 . Wg_in + t -> hid_pre
    dwconv   gate -> gated                                  gemm     gated . Wout + x -> x
  kept per block: xhat, rstd (x2), qkv_pre, qkv, hid_pre, gated, the MDTA workspace (Gram partials, norms, attention).
  The residual stream itself is NOT kept per block (LayerNorm backward works from xhat and rstd).

backward (reverse order; g is the gradient of the residual stream, updated in place):
    FFN   dgated = g . Wout^T            | wgrad(g, gated) -> dWout            | dy = gate_bwd(dwconv(hid_pre), dgated): ONE kernel,
          the stencil is recomputed      | dw_wgrad(hid_pre, dy) -> d dwconv.w | dhid_pre = dwconv(dy, flipped taps)
          dxhat = dhid_pre . Wg_in^T     | wgrad(dhid_pre, xhat2) -> dW_in, dgamma2, dbeta2       | g += ln_bwd(dxhat, xhat2, rstd2)
    MDTA  wgrad per image (g, v) -> dWfold[b]; mdta_bwd: softmax/cosine/temperature backward on the c x c matrices, produces
          dWo, dtemperature and two per-image weight sets so that the pixel-sized work is again GEMMs:
              dv = g . Wfold[b]            d[q|k] = [q|k] . Wqk[b]      (normalisation backward folded into Wqk)
          dw_wgrad(qkv_pre, dqkv), dqkv_pre = dwconv(dqkv, flipped), dxhat = dqkv_pre . Wg_qkv^T, wgrad -> dWqkv, dgamma1, dbeta1,
          g += ln_bwd(dxhat, xhat1, rstd1)
Weight gradients are split-K tcgen05 Grams over the pixels (pir_wgrad) whose fp32 partials are reduced by pir_wgrad_finalize
straight into ONE flat fp32 gradient buffer (`grad_flat`; `grads[name]` are views in parameter order) -- the buffer a
data-parallel step all-reduces with a single NCCL call.

The op records (kind + tensor views + scalars) are interpreted on CPU by tests/emulator.py to check the derivation and the
wiring against autograd of the oracle without a GPU.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ops, packing
from ._lib import LN_BIASFREE, LN_NONE, LN_WITHBIAS, OUT_FINAL_NCHW32, OUT_NHWC16, OUT_SHUFFLE16, OUT_UNSHUFFLE16
from .engine import Engine

Tensor = torch.Tensor


class TrainEngine(Engine):
    """Forward + backward launch programs.  `fwd_ops` / `bwd_ops` are the two record lists."""

    def __init__(self, module, batch: int, height: int, width: int, device, dtype: torch.dtype = torch.bfloat16,
                 grad_scale: Optional[float] = None, input_grad: bool = False):
        # static loss scale: gradients of a mean-reduced loss are ~1/numel and underflow fp16 (not bf16)
        self.grad_scale = float(grad_scale if grad_scale is not None else (65536.0 if dtype == torch.float16 else 1.0))
        self.input_grad = input_grad
        if input_grad and self.grad_scale != 1.0:
            raise NotImplementedError("promptir_b200: the input-image gradient is only available with grad_scale == 1 (bf16)")
        self._graphs: Dict[str, object] = {}
        super().__init__(module, batch, height, width, device, dtype)

    # ------------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------------
    def _f32(self, *shape) -> Tensor:
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)

    def _grad_of(self, p: Optional[Tensor]) -> Optional[Tensor]:
        return None if p is None else self.grads[self._pname[id(p)]]

    def _raw(self, p: Optional[Tensor]) -> Optional[Tensor]:
        """The fp32 parameter itself (finalize kernels read gamma/beta/W from the canonical storage)."""
        if p is None:
            return None
        if self.cuda:
            self.pk._params.append(p)                   # part of the pointer fingerprint (Engine.params_moved)
        return p.detach()

    def _later(self, fn: Callable[[], None]) -> None:
        self._bwd_stack.append(fn)

    def _emit(self, kind: str, launch_fn: Optional[Callable] = None, **args) -> None:
        rec = {"kind": kind, **args}
        rec["launch"] = launch_fn() if (self.cuda and launch_fn is not None and not self._dry) else None
        self.ops.append(rec)

    # ---- record emitters (forward or backward list, whichever is current) ---------------------------------
    def _ln_fwd(self, x, xhat, rstd, tag):
        self._emit("ln_fwd", lambda: ops.ln_fwd(x, xhat, rstd, self.ln_mode), x=x, xhat=xhat, rstd=rstd, ln_mode=self.ln_mode, tag=tag)

    def _ln_bwd(self, d, xhat, rstd, g, tag):
        self._emit("ln_bwd", lambda: ops.ln_bwd(d, xhat, rstd, g, self.ln_mode), d=d, xhat=xhat, rstd=rstd, g=g, ln_mode=self.ln_mode,
                   tag=tag)

    def _dw(self, x, w, out, *, gate, bias, tag, dg=None):
        self._emit("dwconv", lambda: ops.dwconv3x3(x, w, out, gate=gate, bias=bias, dg=dg), x=x, w=w, out=out, gate=gate, bias=bias, dg=dg,
                   tag=tag)

    def _wgrad(self, a, b, *, taps=1, per_image=False, colsum=False, tag=""):
        """partials[img*splits + s][tap][m][n] = sum_{p in split} a[p, m] * b[p + off(tap), n]  (fp32, in self.wg_ws)."""
        B, H, W, M = a.shape
        N = b.shape[3]
        splits = ops.wgrad_splits(B, H * W, M, N, taps, per_image)
        P = B * splits if per_image else splits
        need = P * taps * M * N + (P * M if colsum else 0)
        self._wg_need = max(self._wg_need, need)
        rec = dict(a=a, b=b, taps=taps, per_image=per_image, splits=splits, P=P, M=M, N=N, colsum=colsum, tag=tag)
        self._emit("wgrad", lambda: ops.wgrad(self.wg_ws, rec), ws=self.wg_ws, **rec)
        return self.ops[-1]

    def _wgrad_fin(self, wg: dict, *, dst_w, half=None, half_pad=None, gamma=None, beta=None, w=None, dst_gamma=None, dst_beta=None,
                   dst_bias=None, tag=""):
        R = dst_w.shape[0]
        rec = dict(wg=wg, dst_w=dst_w, half=R if half is None else half, half_pad=R if half_pad is None else half_pad, gamma=gamma,
                   beta=beta, w=w, dst_gamma=dst_gamma, dst_beta=dst_beta, dst_bias=dst_bias, inv_scale=1.0 / self.grad_scale, tag=tag)
        self._emit("wgrad_fin", lambda: ops.wgrad_finalize(self.wg_ws, rec), ws=self.wg_ws, **rec)

    def _dw_wgrad(self, x, dy, *, dst_w, dst_bias, half=None, half_pad=None, tag=""):
        B, H, W, Cp = x.shape
        parts = ops.dw_wgrad_parts(B, H, W, Cp)
        self._wg_need = max(self._wg_need, parts * 10 * Cp)
        R = dst_w.shape[0]
        rec = dict(x=x, dy=dy, parts=parts, dst_w=dst_w, dst_bias=dst_bias, half=R if half is None else half,
                   half_pad=R if half_pad is None else half_pad, inv_scale=1.0 / self.grad_scale, tag=tag)
        self._emit("dw_wgrad", lambda: ops.dw_wgrad(self.wg_ws, rec), ws=self.wg_ws, **rec)

    # ------------------------------------------------------------------------------------------------
    # program construction
    # ------------------------------------------------------------------------------------------------
    def _build(self) -> None:
        m, B, H, W, dt = self.m, self.B, self.H, self.W, self.dtype
        dim = m.patch_embed.proj.out_channels
        size = [(H >> l, W >> l) for l in range(4)]
        self.ln_mode = LN_BIASFREE if m.layernorm_type == "BiasFree" else LN_WITHBIAS
        self._bwd_stack: List[Callable[[], None]] = []
        self._wg_need = 0
        self._dry = False
        self.wg_ws: Optional[Tensor] = None                       # shared fp32 workspace of the weight-gradient kernels
        self.saved_bytes = 0

        # ---- flat gradient buffer, views in parameter order -------------------------------------------
        params = list(m.named_parameters())
        self._pname = {id(p): n for n, p in params}
        total = sum(p.numel() for _, p in params)
        self.grad_flat = self._f32(total)
        self.grads: Dict[str, Tensor] = {}
        off = 0
        for n, p in params:
            self.grads[n] = self.grad_flat[off:off + p.numel()].view(p.shape)
            off += p.numel()

        def blocks_of(mod):
            return list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]

        stage_levels = [(m.encoder_level1, 0), (m.encoder_level2, 1), (m.encoder_level3, 2), (m.latent, 3), (m.noise_level3, 3),
                        (m.decoder_level3, 2), (m.noise_level2, 2), (m.decoder_level2, 1), (m.noise_level1, 1),
                        (m.decoder_level1, 0), (m.refinement, 0)]
        ta = tb = td = 0
        widths = set()
        for mod, lvl in stage_levels:
            for blk in blocks_of(mod):
                c = blk.attn.qkv.in_channels
                hp = packing.round_up(blk.ffn.project_out.in_channels, 8)
                n = B * size[lvl][0] * size[lvl][1]
                ta, tb, td = max(ta, n * hp), max(tb, n * max(3 * c, 2 * hp)), max(td, n * c)
                widths.add(c)
        for pg, lvl in ((m.prompt3, 3), (m.prompt2, 2), (m.prompt1, 1)):
            tb = max(tb, B * size[lvl][0] * size[lvl][1] * pg.conv3x3.in_channels)
        for up, lvl in ((m.up4_3, 3), (m.up3_2, 2), (m.up2_1, 1)):                  # unshuffled gradient of the conv output
            tb = max(tb, B * size[lvl][0] * size[lvl][1] * up.body[0].out_channels)
        for dn, lvl in ((m.down1_2, 0), (m.down2_3, 1), (m.down3_4, 2)):            # shuffled gradient of the conv output
            tb = max(tb, B * size[lvl][0] * size[lvl][1] * dn.body[0].out_channels)
        # backward scratch arenas (16-bit): Ta dgated | Tb y/dy, dqkv, (un)shuffled conv-output gradients | Tc dhid_pre, dqkv_pre | Td dxhat
        self.Ta, self.Tb, self.Tc, self.Td = (torch.zeros(n, dtype=dt, device=self.device) for n in (ta, tb, tb, td))
        self.wfold: Dict[int, Tensor] = {c: self._zeros(B, c, packing.kpad_of(c)) for c in sorted(widths)}
        self.wft: Dict[int, Tensor] = {c: self._zeros(B, c, packing.kpad_of(c)) for c in sorted(widths)}
        self.wqk: Dict[int, Tensor] = {c: self._zeros(B, 2 * c, packing.kpad_of(2 * c)) for c in sorted(widths)}

        c1, c2, c3, c4 = dim, dim * 2, dim * 4, dim * 8
        p1, p2, p3 = (pg.conv3x3.in_channels for pg in (m.prompt1, m.prompt2, m.prompt3))
        (h0, w0), (h1, w1), (h2, w2), (h3, w3) = size
        up1 = m.up2_1.body[0].out_channels // 4
        up2 = m.up3_2.body[0].out_channels // 4
        up3 = m.up4_3.body[0].out_channels // 4
        # activations (same concat folding as the inference engine) and their gradients
        shapes = dict(cat1=(h0, w0, up1 + c1), cat2=(h1, w1, up2 + c2), cat3=(h2, w2, up3 + c3), catn3=(h3, w3, c4 + p3),
                      catn2=(h2, w2, c3 + p2), catn1=(h1, w1, c2 + p1), r3=(h3, w3, m.reduce_noise_level3.out_channels),
                      r2=(h2, w2, m.reduce_noise_level2.out_channels), r1=(h1, w1, m.reduce_noise_level1.out_channels),
                      dec1=(h0, w0, up1 + c1))
        for name, shp in shapes.items():
            setattr(self, name, self._zeros(B, *shp))
            if name != "dec1":
                setattr(self, "g_" + name, self._zeros(B, *shp))
        self.g_cat1 = None                                   # the gradient of cat1 is the (in-place) gradient of the dec1 stream
        self.g_dec1 = self._zeros(B, *shapes["dec1"])
        cin, cout = m.patch_embed.proj.in_channels, m.output.out_channels
        self.img_in = self._f32(B, cin, H, W)
        self.out = self._f32(B, cout, H, W)
        self.d_out = self._f32(B, cout, H, W)                # dL/d(out), fp32 NCHW (what autograd hands to backward)
        self.d_out8 = self._zeros(B, H, W, 8)                # the same as NHWC 16-bit, channels padded to 8, times grad_scale
        self.img8 = self._zeros(B, H, W, 8)                  # the input image as NHWC 16-bit (patch-embed wgrad operand)
        self.d_img = self._f32(B, cin, H, W) if self.input_grad else None
        assert cin <= 8 and cout <= 8

        enc1, enc2, enc3 = self.cat1[..., up1:], self.cat2[..., up2:], self.cat3[..., up3:]
        lat, d3, d2 = self.catn3[..., :c4], self.catn2[..., :c3], self.catn1[..., :c2]
        g_enc1, g_enc2, g_enc3 = self.g_dec1[..., up1:], self.g_cat2[..., up2:], self.g_cat3[..., up3:]
        g_lat, g_d3, g_d2 = self.g_catn3[..., :c4], self.g_catn2[..., :c3], self.g_catn1[..., :c2]

        # ================================ forward program ================================================
        self.ops = self.fwd_ops = []
        pe = m.patch_embed.proj
        pe_w, pe_b = self.pk.f32(pe.weight), self.pk.f32(pe.bias)
        self._emit("patch_embed", lambda: ops.patch_embed(self.img_in, pe_w, pe_b, enc1), img=self.img_in, w=pe_w, bias=pe_b, out=enc1)
        self._later(lambda: self._patch_embed_bwd(pe, g_enc1))
        self._tstage(m.encoder_level1, enc1, g_enc1)
        self._tdown(m.down1_2, enc1, enc2, g_enc1, g_enc2)
        self._tstage(m.encoder_level2, enc2, g_enc2)
        self._tdown(m.down2_3, enc2, enc3, g_enc2, g_enc3)
        self._tstage(m.encoder_level3, enc3, g_enc3)
        self._tdown(m.down3_4, enc3, lat, g_enc3, g_lat)
        self._tstage(m.latent, lat, g_lat)
        self._tprompt(m.prompt3, lat, self.catn3[..., c4:], g_lat, self.g_catn3[..., c4:])
        self._tstage(m.noise_level3, self.catn3, self.g_catn3)
        self._treduce(m.reduce_noise_level3, self.catn3, self.r3, self.g_catn3, self.g_r3)
        self._tup(m.up4_3, self.r3, self.cat3[..., :up3], self.g_r3, self.g_cat3[..., :up3])
        self._treduce(m.reduce_chan_level3, self.cat3, d3, self.g_cat3, g_d3)
        self._tstage(m.decoder_level3, d3, g_d3)
        self._tprompt(m.prompt2, d3, self.catn2[..., c3:], g_d3, self.g_catn2[..., c3:])
        self._tstage(m.noise_level2, self.catn2, self.g_catn2)
        self._treduce(m.reduce_noise_level2, self.catn2, self.r2, self.g_catn2, self.g_r2)
        self._tup(m.up3_2, self.r2, self.cat2[..., :up2], self.g_r2, self.g_cat2[..., :up2])
        self._treduce(m.reduce_chan_level2, self.cat2, d2, self.g_cat2, g_d2)
        self._tstage(m.decoder_level2, d2, g_d2)
        self._tprompt(m.prompt1, d2, self.catn1[..., c2:], g_d2, self.g_catn1[..., c2:])
        self._tstage(m.noise_level1, self.catn1, self.g_catn1)
        self._treduce(m.reduce_noise_level1, self.catn1, self.r1, self.g_catn1, self.g_r1)
        self._tup(m.up2_1, self.r1, self.cat1[..., :up1], self.g_r1, self.g_dec1[..., :up1])
        # decoder_level1 would overwrite cat1 (whose encoder half down1_2's weight gradient still needs): its first block
        # writes the stream to dec1 instead; the gradient of dec1 flows back into cat1's gradient through the same buffer
        self._tstage(m.decoder_level1, self.cat1, self.g_dec1, first_out=self.dec1)
        self._tstage(m.refinement, self.dec1, self.g_dec1)
        oc = m.output
        ow, ob = self.pk.conv3x3(oc.weight, dt), self.pk.f32(oc.bias)
        self._gemm(self.dec1, ow, self.out, n=oc.out_channels, taps=9, out_mode=OUT_FINAL_NCHW32, vec_t=ob, img=self.img_in, tag="output")
        self._later(lambda: self._output_bwd(oc))

        self._finish_build()

    def _finish_build(self) -> None:
        """Emit the backward program from the closures the forward construction left on the stack (in reverse), after a dry pass that
        sizes the shared fp32 workspace of the weight-gradient kernels."""
        npk, mark = len(self._packers), self.pk.mark()
        self.ops, self._dry = [], True
        for fn in reversed(self._bwd_stack):
            fn()
        del self._packers[npk:]
        self.pk.rollback(mark)                          # the dry pass's cache requests are made again by the real pass
        self.wg_ws = self._f32(max(self._wg_need, 1))
        self.ops, self._dry = [], False
        self.bwd_ops = self.ops
        for fn in reversed(self._bwd_stack):
            fn()
        self._bwd_stack = []
        self.ops = self.fwd_ops + self.bwd_ops
        # parameters that receive a gradient = those some backward record writes (the reference's 6 dead convs get none)
        written = set()
        for r in self.bwd_ops:
            for k, v in r.items():
                if k.startswith("dst_") and v is not None:
                    written.add(v.data_ptr())
        self.live_params = {n for n, gview in self.grads.items() if gview.data_ptr() in written}
        self.generation = 0
        self._param_version = self._current_version()
        self.fwd_launches = [r["launch"] for r in self.fwd_ops]
        self.bwd_launches = [r["launch"] for r in self.bwd_ops]
        self.launches = self.fwd_launches

    # ---- stages ------------------------------------------------------------------------------------------
    def _tstage(self, mod, x: Tensor, g: Tensor, first_out: Optional[Tensor] = None) -> None:
        blocks = list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]
        for i, blk in enumerate(blocks):
            if i == 0 and first_out is not None:
                self._tblock(blk, x, g, x_out=first_out)
                x = first_out
            else:
                self._tblock(blk, x, g)

    def _tblock(self, blk, x: Tensor, g: Tensor, x_out: Optional[Tensor] = None) -> None:
        """One TransformerBlock, forward now and its backward pushed on the stack.  net/model.py:192-196."""
        xo = x if x_out is None else x_out
        self._t_mdta(blk.attn, blk.norm1, x, g, xo)
        self._t_gdfn(blk.ffn, blk.norm2, xo, g)

    def _t_mdta(self, at, norm, x: Tensor, g: Tensor, xo: Tensor) -> None:
        """xo = x + MDTA(LN(x)) (net/model.py:117-138, 193) and its backward on g (in place)."""
        dt = self.dtype
        B, h, w, c = x.shape
        heads = at.num_heads
        n1 = norm.body
        beta1 = getattr(n1, "bias", None)
        flip = lambda wt: wt.detach().flip(2, 3)
        pk = self.pk
        qkv_w, _, qkv_t = pk.pointwise(at.qkv.weight, dt, gamma=n1.weight, beta=beta1, bias=at.qkv.bias)
        qkv_wT = pk.pointwise(at.qkv.weight, dt, gamma=n1.weight, transpose=True)[0]            # (W diag(gamma))^T for the dgrad GEMM
        dwq_w, dwq_f = pk.depthwise(at.qkv_dwconv.weight, dt), pk.depthwise(at.qkv_dwconv.weight, dt, flip=True)
        dwq_b = pk.f32(at.qkv_dwconv.bias)
        temp, wo, wo_b = pk.f32(at.temperature, (-1,)), pk.f32(at.project_out.weight, (c, c)), pk.f32(at.project_out.bias)
        keep = lambda ch: self._zeros(B, h, w, ch)
        xh1, qkv_pre, qkv = keep(c), keep(3 * c), keep(3 * c)
        rstd1 = self._f32(B * h * w)
        splits = ops.mdta_splits(B, h * w, c)
        ws = self._f32(ops.mdta_ws_floats(B, c, splits))
        self.saved_bytes += sum(t.numel() * t.element_size() for t in (xh1, qkv_pre, qkv, rstd1, ws))
        wfold = self.wfold[c]

        self._ln_fwd(x, xh1, rstd1, "LN1")
        self._gemm(xh1, qkv_w, qkv_pre, n=3 * c, vec_t=qkv_t, tag="K1")
        self._dw(qkv_pre, dwq_w, qkv, gate=False, bias=dwq_b, tag="K2")
        gram_fin = ops.mdta(qkv, heads, ws, temp, wo, wfold, splits) if self.cuda else (None, None)
        self._emit("mdta_gram", (lambda: gram_fin[0]), qkv=qkv, heads=heads, ws=ws, splits=splits, tag="K3a")
        self._emit("mdta_finalize", (lambda: gram_fin[1]), qkv=qkv, heads=heads, ws=ws, splits=splits, temperature=temp, wo=wo,
                   wfold=wfold, tag="K3b")
        self._gemm(qkv[..., 2 * c:], wfold, xo, n=c, res=x, vec_t=wo_b, w_batched=True, tag="K4")

        def bwd():
            G = self._grad_of
            dxh = self._scratch(self.Td, h, w, c)
            dqkv = self._scratch(self.Tb, h, w, 3 * c)
            dqkv_pre = self._scratch(self.Tc, h, w, 3 * c)
            v = qkv[..., 2 * c:]
            wgv = self._wgrad(g, v, per_image=True, colsum=at.project_out.bias is not None, tag="B4w")
            wft, wqk = self.wft[c], self.wqk[c]
            rec = dict(wg=wgv, fws=ws, qkv=qkv, splits=splits, heads=heads, B=B, C=c, HW=h * w, temperature=self._raw(at.temperature), wo=wo,
                       wft=wft, wqk=wqk, dst_wo=G(at.project_out.weight).view(c, c), dst_temp=G(at.temperature).view(-1),
                       dst_bias=G(at.project_out.bias), inv_scale=1.0 / self.grad_scale, tag="B3")
            self._wg_need = max(self._wg_need, wgv["P"] * (c * c + c) + ops.mdta_bwd_ws_floats(B, c, heads))
            self._emit("mdta_bwd", lambda: ops.mdta_bwd(self.wg_ws, rec), ws=self.wg_ws, **rec)
            self._gemm(g, wft, dqkv[..., 2 * c:], n=c, w_batched=True, tag="B4d")
            self._gemm(qkv[..., :2 * c], wqk, dqkv[..., :2 * c], n=2 * c, w_batched=True, tag="B3d")
            self._dw_wgrad(qkv_pre, dqkv, dst_w=G(at.qkv_dwconv.weight), dst_bias=G(at.qkv_dwconv.bias), tag="B2w")
            self._dw(dqkv, dwq_f, dqkv_pre, gate=False, bias=None, tag="B2d")
            self._gemm(dqkv_pre, qkv_wT, dxh, n=c, tag="B1d")
            wg = self._wgrad(dqkv_pre, xh1, colsum=True, tag="B1w")
            self._wgrad_fin(wg, dst_w=G(at.qkv.weight).view(3 * c, c), gamma=self._raw(n1.weight), beta=self._raw(beta1),
                            w=self._raw(at.qkv.weight), dst_gamma=G(n1.weight), dst_beta=G(beta1), dst_bias=G(at.qkv.bias), tag="B1f")
            self._ln_bwd(dxh, xh1, rstd1, g, "BLN1")
        self._later(bwd)

    def _t_gdfn(self, ff, norm, x: Tensor, g: Tensor) -> None:
        """x += GDFN(LN(x)) in place (net/model.py:94-99, 194) and its backward on g (in place)."""
        dt = self.dtype
        B, h, w, c = x.shape
        hid = ff.project_out.in_channels
        hp, gmap = packing.gdfn_maps(hid, self.device)
        n2 = norm.body
        beta2 = getattr(n2, "bias", None)
        flip = lambda wt: wt.detach().flip(2, 3)
        pk = self.pk
        pin_w, _, pin_t = pk.pointwise(ff.project_in.weight, dt, gamma=n2.weight, beta=beta2, bias=ff.project_in.bias, rows=(hid, hp),
                                       n_total=2 * hp)
        pin_wT = pk.pointwise(ff.project_in.weight, dt, gamma=n2.weight, transpose=True, cols=(hid, hp), k_total=2 * hp)[0]
        dwf_w = pk.depthwise(ff.dwconv.weight, dt, split=(hid, hp), c_total=2 * hp)
        dwf_f = pk.depthwise(ff.dwconv.weight, dt, flip=True, split=(hid, hp), c_total=2 * hp)
        dwf_b = pk.vec(ff.dwconv.bias, (hid, hp), 2 * hp)
        pout_w, _, pout_t = pk.pointwise(ff.project_out.weight, dt, bias=ff.project_out.bias, k_total=hp)
        pout_wT = pk.pointwise(ff.project_out.weight, dt, transpose=True, n_total=hp)[0]
        keep = lambda ch: self._zeros(B, h, w, ch)
        xh2, hid_pre, gated = keep(c), keep(2 * hp), keep(hp)
        rstd2 = self._f32(B * h * w)
        self.saved_bytes += sum(t.numel() * t.element_size() for t in (xh2, hid_pre, gated, rstd2))

        self._ln_fwd(x, xh2, rstd2, "LN2")
        self._gemm(xh2, pin_w, hid_pre, n=2 * hp, vec_t=pin_t, tag="K5")
        self._dw(hid_pre, dwf_w, gated, gate=True, bias=dwf_b, tag="K6")
        self._gemm(gated, pout_w, x, n=c, res=x, vec_t=pout_t, tag="K7")

        def bwd():
            G = self._grad_of
            dgt = self._scratch(self.Ta, h, w, hp)
            y = self._scratch(self.Tb, h, w, 2 * hp)
            dhp = self._scratch(self.Tc, h, w, 2 * hp)
            dxh = self._scratch(self.Td, h, w, c)
            self._gemm(g, pout_wT, dgt, n=hp, tag="B7d")
            wg = self._wgrad(g, gated, colsum=ff.project_out.bias is not None, tag="B7w")
            self._wgrad_fin(wg, dst_w=G(ff.project_out.weight).view(c, hid), dst_bias=G(ff.project_out.bias), tag="B7f")
            self._dw(hid_pre, dwf_w, y, gate=2, bias=dwf_b, dg=dgt, tag="B6g")      # y = dw(hid_pre) recomputed, gate backward fused
            self._dw_wgrad(hid_pre, y, dst_w=G(ff.dwconv.weight), dst_bias=G(ff.dwconv.bias), half=hid, half_pad=hp, tag="B6w")
            self._dw(y, dwf_f, dhp, gate=False, bias=None, tag="B6d")
            self._gemm(dhp, pin_wT, dxh, n=c, tag="B5d")
            wg = self._wgrad(dhp, xh2, colsum=True, tag="B5w")
            self._wgrad_fin(wg, dst_w=G(ff.project_in.weight).view(2 * hid, c), half=hid, half_pad=hp, gamma=self._raw(n2.weight),
                            beta=self._raw(beta2), w=self._raw(ff.project_in.weight), dst_gamma=G(n2.weight), dst_beta=G(beta2),
                            dst_bias=G(ff.project_in.bias), tag="B5f")
            self._ln_bwd(dxh, xh2, rstd2, g, "BLN2")
        self._later(bwd)

    # ---- down / up / reduce / prompt / ends --------------------------------------------------------------------
    def _tdown(self, mod, x, out, g_x, g_out) -> None:
        conv = mod.body[0]                                   # model.py:164-165: conv3x3 n -> n/2, PixelUnshuffle(2)
        dt = self.dtype
        w, wT = self.pk.conv3x3(conv.weight, dt), self.pk.conv3x3(conv.weight, dt, transpose_flip=True)
        self._gemm(x, w, out, n=conv.out_channels, taps=9, out_mode=OUT_UNSHUFFLE16, tag="down")
        B, h, ww, _ = x.shape

        def bwd():
            dconv = self._scratch(self.Tb, h, ww, conv.out_channels)
            self._emit("shuffle", lambda: ops.pixel_shuffle(g_out, dconv, up=True), x=g_out, out=dconv, up=True, tag="Bsh")
            wg = self._wgrad(dconv, x, taps=9, tag="Bdw")
            self._wgrad_fin(wg, dst_w=self._grad_of(conv.weight), tag="Bdf")
            self._gemm(dconv, wT, g_x, n=conv.in_channels, taps=9, res=g_x, tag="Bdd")      # the skip gradient is already in g_x
        self._later(bwd)

    def _tup(self, mod, x, out, g_x, g_out) -> None:
        conv = mod.body[0]                                   # model.py:174-175: conv3x3 n -> 2n, PixelShuffle(2)
        dt = self.dtype
        w, wT = self.pk.conv3x3(conv.weight, dt), self.pk.conv3x3(conv.weight, dt, transpose_flip=True)
        self._gemm(x, w, out, n=conv.out_channels, taps=9, out_mode=OUT_SHUFFLE16, tag="up")
        B, h, ww, _ = x.shape

        def bwd():
            dconv = self._scratch(self.Tb, h, ww, conv.out_channels)
            self._emit("shuffle", lambda: ops.pixel_shuffle(g_out, dconv, up=False), x=g_out, out=dconv, up=False, tag="Bus")
            wg = self._wgrad(dconv, x, taps=9, tag="Buw")
            self._wgrad_fin(wg, dst_w=self._grad_of(conv.weight), tag="Buf")
            self._gemm(dconv, wT, g_x, n=conv.in_channels, taps=9, tag="Bud")
        self._later(bwd)

    def _tconv3(self, conv, x, out, g_x, g_out, tag="conv3") -> None:
        """out = conv3x3(x) (no bias), g_x = conv3x3^T(g_out) (assigned)."""
        dt = self.dtype
        w, wT = self.pk.conv3x3(conv.weight, dt), self.pk.conv3x3(conv.weight, dt, transpose_flip=True)
        self._gemm(x, w, out, n=conv.out_channels, taps=9, tag=tag)

        def bwd():
            wg = self._wgrad(g_out, x, taps=9, tag="B" + tag + "w")
            self._wgrad_fin(wg, dst_w=self._grad_of(conv.weight), tag="B" + tag + "f")
            self._gemm(g_out, wT, g_x, n=conv.in_channels, taps=9, tag="B" + tag + "d")
        self._later(bwd)

    def _treduce(self, conv, x, out, g_x, g_out) -> None:
        dt = self.dtype
        w, _, t = self.pk.pointwise(conv.weight, dt, bias=conv.bias)
        wT = self.pk.pointwise(conv.weight, dt, transpose=True)[0]
        self._gemm(x, w, out, n=conv.out_channels, vec_t=t, tag="reduce")

        def bwd():
            wg = self._wgrad(g_out, x, colsum=conv.bias is not None, tag="Brw")
            self._wgrad_fin(wg, dst_w=self._grad_of(conv.weight).view(conv.out_channels, conv.in_channels), dst_bias=self._grad_of(conv.bias),
                            tag="Brf")
            self._gemm(g_out, wT, g_x, n=conv.in_channels, tag="Brd")
        self._later(bwd)

    def _tprompt(self, pg, x, out, g_x, g_out, align_corners: bool = False) -> None:
        """PromptGenBlock (model.py:226-235).  g_x receives (+=) the gradient through the global-average-pool branch."""
        dt = self.dtype
        B, h, w, cx = x.shape
        d = pg.conv3x3.in_channels
        L = pg.prompt_param.shape[1]
        prm, lw, lb = self.pk.prompt(pg.prompt_param), self.pk.f32(pg.linear_layer.weight), self.pk.f32(pg.linear_layer.bias)
        cw, cwT = self.pk.conv3x3(pg.conv3x3.weight, dt), self.pk.conv3x3(pg.conv3x3.weight, dt, transpose_flip=True)
        up = self._zeros(B, h, w, d)                          # kept: resized prompt (operand of the conv's weight gradient)
        pws = self._f32(ops.prompt_ws_floats(B, h * w, cx))   # kept: pooled partial sums
        wts = self._f32(B, L)                                 # kept: softmax weights
        self._emit("prompt", lambda: ops.prompt_gen(x, prm, lw, lb, up, pws, wts, align_corners=align_corners), x=x, prompt=prm, lin_w=lw,
                   lin_b=lb, out=up, ws=pws, weights_out=wts, align_corners=align_corners, tag="K10")
        self._gemm(up, cw, out, n=d, taps=9, tag="prompt_conv")

        def bwd():
            dup = self._scratch(self.Tb, h, w, d)
            wg = self._wgrad(g_out, up, taps=9, tag="Bpw")
            self._wgrad_fin(wg, dst_w=self._grad_of(pg.conv3x3.weight), tag="Bpf")
            self._gemm(g_out, cwT, dup, n=d, taps=9, tag="Bpd")
            demb = self._f32(B, cx)
            rec = dict(dup=dup, prompt=prm, weights=wts, pool_ws=pws, lin_w=lw, HW=h * w, C=cx, demb=demb, align_corners=align_corners,
                       dst_prompt=self._grad_of(pg.prompt_param),
                       dst_lin_w=self._grad_of(pg.linear_layer.weight), dst_lin_b=self._grad_of(pg.linear_layer.bias),
                       inv_scale=1.0 / self.grad_scale, tag="Bpp")
            self._wg_need = max(self._wg_need, ops.prompt_bwd_ws_floats(B, L, d, prm.shape[1]))
            self._emit("prompt_bwd", lambda: ops.prompt_bwd(self.wg_ws, rec), ws=self.wg_ws, **rec)
            self._emit("bcast_add", lambda: ops.bcast_add(g_x, demb), g=g_x, v=demb, tag="Bpb")
        self._later(bwd)

    def _output_bwd(self, oc) -> None:
        """out = conv3x3(dec1) + img (model.py:377).  First backward op: dL/dout arrives as fp32 NCHW in self.d_out."""
        dt = self.dtype
        wT = self.pk.conv3x3(oc.weight, dt, transpose_flip=True)
        self._emit("to_nhwc16", lambda: ops.nchw32_to_nhwc16(self.d_out, self.d_out8, self.grad_scale), src=self.d_out, out=self.d_out8,
                   scale=self.grad_scale, tag="Bo8")
        wg = self._wgrad(self.d_out8, self.dec1, taps=9, colsum=oc.bias is not None, tag="Bow")
        self._wgrad_fin(wg, dst_w=self._grad_of(oc.weight), dst_bias=self._grad_of(oc.bias), tag="Bof")
        self._gemm(self.d_out8, wT, self.g_dec1, n=oc.in_channels, taps=9, tag="Bod")

    def _patch_embed_bwd(self, pe, g_enc1) -> None:
        """Last backward op(s): weight gradient of the 3x3 patch embedding (model.py:206) and, on request, dL/d(input)."""
        self._emit("to_nhwc16", lambda: ops.nchw32_to_nhwc16(self.img_in, self.img8, 1.0), src=self.img_in, out=self.img8, scale=1.0, tag="Bi8")
        wg = self._wgrad(g_enc1, self.img8, taps=9, colsum=pe.bias is not None, tag="Bew")
        self._wgrad_fin(wg, dst_w=self._grad_of(pe.weight), dst_bias=self._grad_of(pe.bias), tag="Bef")
        if self.input_grad:                                  # d img = d out (through "+ inp_img") + conv^T(g)
            wT = self.pk.conv3x3(pe.weight, self.dtype, transpose_flip=True)
            self._gemm(g_enc1, wT, self.d_img, n=pe.in_channels, taps=9, out_mode=OUT_FINAL_NCHW32, img=self.d_out, tag="Bed")

    # ------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------
    def _run(self, which: str, launches, use_graph: bool) -> None:
        stream = torch.cuda.current_stream(self.device)
        if not use_graph:
            for fn in launches:
                fn(stream.cuda_stream)
            return
        g = self._graphs.get(which)
        if g is None:
            for fn in launches:                              # warm: kernel attributes are set outside of capture
                fn(stream.cuda_stream)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s = torch.cuda.current_stream(self.device).cuda_stream
                for fn in launches:
                    fn(s)
            self._graphs[which] = g
        g.replay()

    def forward(self, img: Tensor, use_graph: bool = True) -> Tensor:
        """Training forward: keeps activations for `backward`.  Returns a fresh fp32 NCHW tensor."""
        if not self.cuda:
            raise RuntimeError("promptir_b200.TrainEngine needs a CUDA (sm_100a) device; there is no CPU path")
        if tuple(img.shape) != tuple(self.img_in.shape):
            raise ValueError(f"engine built for {tuple(self.img_in.shape)}, got {tuple(img.shape)}")
        if self._current_version() != self._param_version:
            self.refresh_weights()
        self.img_in.copy_(img)
        self._run("fwd", self.fwd_launches, use_graph)
        return self.out.clone()

    def backward(self, d_out: Tensor, use_graph: bool = True) -> None:
        """d_out = dL/d(output) (fp32 NCHW).  Fills `grad_flat` / `grads` (and `d_img` when input_grad)."""
        if not self.cuda:
            raise RuntimeError("promptir_b200.TrainEngine needs a CUDA (sm_100a) device; there is no CPU path")
        self.d_out.copy_(d_out)
        self._run("bwd", self.bwd_launches, use_graph)
        # fp16 storage runs the backward under a static loss scale (65536): an upstream gradient of order 1 overflows the 16-bit range.
        # One reduction over the flat buffer leaves a device-side flag (no sync here); overflowed() reads it.
        # (the 16-bit conversions saturate at +-65504, so an overflow shows up as a clipped dL/d(out) rather than as inf)
        self._found_inf = None
        if self.grad_scale != 1.0:
            self._found_inf = (~torch.isfinite(self.grad_flat)).any() | (self.d_out.abs().amax() * self.grad_scale > 65504.0)

    def overflowed(self) -> bool:
        """True when the last backward produced a non-finite gradient (fp16 loss scale exceeded): skip the optimizer step and lower
        `module.grad_scale`, as torch.cuda.amp.GradScaler would.  Always False for bf16 (scale 1).  Synchronises."""
        f = getattr(self, "_found_inf", None)
        return bool(f.item()) if f is not None else False

    def run_bwd_range(self, a: int, b: int, use_graph: bool = True) -> None:
        """Records [a, b) of the backward program (ddp.OverlappedReducer cuts the backward into a few such pieces); d_out is already set."""
        self._run(("bwd", a, b), self.bwd_launches[a:b], use_graph)

    def replay(self, use_graph: bool = True) -> None:
        self._run("fwd", self.fwd_launches, use_graph)

    def kernels_per_step(self) -> int:
        per = {"mdta_finalize": 2, "prompt": 2, "wgrad_fin": 2, "dw_wgrad": 2, "mdta_bwd": 6, "prompt_bwd": 4}
        return sum(getattr(r.get("launch"), "kernels", per.get(r["kind"], 1)) for r in self.ops)

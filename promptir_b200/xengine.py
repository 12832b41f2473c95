"""Execution engine of the PromptXRestormer forward (net/prompt_xrestormer.py:428-478) on one B200.

Same machinery as `engine.Engine` (NHWC 16-bit arena, concat folding, packed-weight cache, prepared C-ABI launches replayed as a
CUDA graph); only the program differs.  Per X-TransformerBlock (prompt_xrestormer.py:255-260):

    channel attention   K12 pwdw  LN1 + 1x1 (C -> 3C) + dw3x3 | K3 mdta_gram + finalize | K4 gemm v.Wfold[b] + residual
    channel FFN         K56 pwdw  LN2 + 1x1 + dw3x3 + GELU gate | K7 gemm + residual
    spatial attention   gemm LN3-folded 1x1 (C -> 3*inner) | ocab (window attention, relative position bias) | gemm inner -> C + residual
    spatial FFN         K56 pwdw  LN4 + ... | K7 gemm + residual

(the unfused K1/K2, K5/K6 pairs where the x tile does not fit shared memory: the 160/320/704-wide prompt blocks).
PromptBlock (prompt_xrestormer.py:343-359): prompt_gen with align_corners=True -> conv3x3 into the right slice of the concat
buffer whose left slice the previous stage already wrote -> one X block on the concatenation -> conv3x3 back to lin_dim.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops, packing
from ._lib import LN_BIASFREE, LN_WITHBIAS, OUT_FINAL_NCHW32
from .engine import Engine

Tensor = torch.Tensor


class XEngine(Engine):
    def _build(self) -> None:
        m, B, H, W, dt = self.m, self.B, self.H, self.W, self.dtype
        if H % 64 or W % 64:
            raise ValueError("PromptXRestormer needs height and width that are multiples of 64")
        dim = m.patch_embed.proj.out_channels
        size = [(H >> l, W >> l) for l in range(4)]
        self.ln_mode = LN_BIASFREE if m.layernorm_type == "BiasFree" else LN_WITHBIAS

        def blocks_of(mod):
            return list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]

        stage_levels = [(m.encoder_level1, 0), (m.encoder_level2, 1), (m.encoder_level3, 2), (m.latent, 3), (m.prompt3.attn, 3),
                        (m.decoder_level3, 2), (m.prompt2.attn, 2), (m.decoder_level2, 1), (m.prompt1.attn, 1), (m.decoder_level1, 0),
                        (m.refinement, 0)]
        s1 = s2 = ws_f = 0
        widths = set()
        for mod, lvl in stage_levels:
            for blk in blocks_of(mod):
                c = blk.channel_attn.qkv.in_channels
                hp = packing.round_up(blk.channel_ffn.project_out.in_channels, 8)
                inner = blk.spatial_attn.inner_dim
                hw = size[lvl][0] * size[lvl][1]
                s1 = max(s1, B * hw * max(3 * c, 2 * hp, 3 * inner))
                s2 = max(s2, B * hw * max(3 * c, hp, inner))
                sp = ops.mdta_splits(B, hw, c)
                ws_f = max(ws_f, ops.mdta_ws_floats(B, c, sp))
                widths.add(c)
        for pg, lvl in ((m.prompt3, 3), (m.prompt2, 2), (m.prompt1, 1)):
            hw = size[lvl][0] * size[lvl][1]
            s2 = max(s2, B * hw * pg.conv3x3.in_channels)
            ws_f = max(ws_f, ops.prompt_ws_floats(B, hw, pg.linear_layer.in_features))
        self.S1 = torch.empty(s1, dtype=dt, device=self.device)
        self.S2 = torch.empty(s2, dtype=dt, device=self.device)
        self.ws = torch.empty(ws_f, dtype=torch.float32, device=self.device)
        self.wfold: Dict[int, Tensor] = {c: self._zeros(B, c, packing.kpad_of(c)) for c in sorted(widths)}

        c1, c2, c3, c4 = dim, dim * 2, dim * 4, dim * 8
        p1, p2, p3 = (pg.conv3x3.in_channels for pg in (m.prompt1, m.prompt2, m.prompt3))
        (h0, w0), (h1, w1), (h2, w2), (h3, w3) = size
        up1 = m.up2_1.body[0].out_channels // 4
        up2 = m.up3_2.body[0].out_channels // 4
        up3 = m.up4_3.body[0].out_channels // 4
        self.cat1 = self._zeros(B, h0, w0, up1 + c1)       # [up2_1 | encoder_level1]          prompt_xrestormer.py:470
        self.cat2 = self._zeros(B, h1, w1, up2 + c2)       # [up3_2 | encoder_level2]          :462
        self.cat3 = self._zeros(B, h2, w2, up3 + c3)       # [up4_3 | encoder_level3]          :454
        self.pcat3 = self._zeros(B, h3, w3, c4 + p3)       # [latent | prompt]  (PromptBlock)  :354
        self.pcat2 = self._zeros(B, h2, w2, c3 + p2)       # [decoder_level3 | prompt]
        self.pcat1 = self._zeros(B, h1, w1, c2 + p1)       # [decoder_level2 | prompt]
        self.r3 = self._zeros(B, h3, w3, c4)
        self.r2 = self._zeros(B, h2, w2, c3)
        self.r1 = self._zeros(B, h1, w1, c2)
        cin, cout = m.patch_embed.proj.in_channels, m.output.out_channels
        self.img_in = self._io[0] if self._io[0] is not None else torch.zeros(B, cin, H, W, dtype=torch.float32, device=self.device)
        self.out = self._io[1] if self._io[1] is not None else torch.zeros(B, cout, H, W, dtype=torch.float32, device=self.device)

        enc1, enc2, enc3 = self.cat1[..., up1:], self.cat2[..., up2:], self.cat3[..., up3:]
        lat, d3, d2 = self.pcat3[..., :c4], self.pcat2[..., :c3], self.pcat1[..., :c2]

        pe = m.patch_embed.proj
        pe_w, pe_b = self._cached(lambda: [pe.weight.detach().float().contiguous(),
                                           None if pe.bias is None else pe.bias.detach().float().contiguous()])
        self._emit("patch_embed", lambda: ops.patch_embed(self.img_in, pe_w, pe_b, enc1), img=self.img_in, w=pe_w, bias=pe_b, out=enc1)
        self._xstage(m.encoder_level1, enc1)
        self._down(m.down1_2, enc1, enc2)
        self._xstage(m.encoder_level2, enc2)
        self._down(m.down2_3, enc2, enc3)
        self._xstage(m.encoder_level3, enc3)
        self._down(m.down3_4, enc3, lat)
        self._xstage(m.latent, lat)
        self._xprompt(m.prompt3, self.pcat3, c4, self.r3)
        self._up(m.up4_3, self.r3, self.cat3[..., :up3])
        self._reduce(m.reduce_chan_level3, self.cat3, d3)
        self._xstage(m.decoder_level3, d3)
        self._xprompt(m.prompt2, self.pcat2, c3, self.r2)
        self._up(m.up3_2, self.r2, self.cat2[..., :up2])
        self._reduce(m.reduce_chan_level2, self.cat2, d2)
        self._xstage(m.decoder_level2, d2)
        self._xprompt(m.prompt1, self.pcat1, c2, self.r1)
        self._up(m.up2_1, self.r1, self.cat1[..., :up1])
        self._xstage(m.decoder_level1, self.cat1)
        self._xstage(m.refinement, self.cat1)
        oc = m.output
        (ow,) = self._cached(lambda: [packing.pack_conv3x3(oc.weight, dt)])
        (ob,) = self._cached(lambda: [None if oc.bias is None else oc.bias.detach().float().contiguous()])
        self._gemm(self.cat1, ow, self.out, n=oc.out_channels, taps=9, out_mode=OUT_FINAL_NCHW32, vec_t=ob, img=self.img_in, tag="output")
        self._param_version = self._current_version()
        self.launches = [r["launch"] for r in self.ops]

    def _xstage(self, mod, x: Tensor) -> None:
        for blk in (list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]):
            self._xblock(blk, x)

    def _ffn(self, ff, norm, x: Tensor) -> None:
        """x += GDFN(LN(x))   (prompt_xrestormer.py:133-152).  Same kernels as net/model.py's FeedForward."""
        dt = self.dtype
        B, h, w, c = x.shape
        hid = ff.project_out.in_channels
        hp, gmap = packing.gdfn_maps(hid, self.device)
        n2 = norm.body
        beta = getattr(n2, "bias", None)
        pin_w, pin_s, pin_t = self._cached(lambda: list(packing.pack_pointwise(
            ff.project_in.weight, dt, gamma=n2.weight, beta=beta, bias=ff.project_in.bias, row_map=gmap, n_total=2 * hp)))
        dwf_w, dwf_b = self._cached(lambda: [packing.pack_depthwise(ff.dwconv.weight, dt, chan_map=gmap, c_total=2 * hp),
                                             packing.scatter_vec(ff.dwconv.bias, gmap, 2 * hp)])
        pout_w, _, pout_t = self._cached(lambda: list(packing.pack_pointwise(ff.project_out.weight, dt, bias=ff.project_out.bias, k_total=hp)))
        hid_pre = self._scratch(self.S1, h, w, 2 * hp)
        gated = self._scratch(self.S2, h, w, hp)
        if self.fuse and ops.pwdw_supported(c, hp, True):
            (dwf_h,) = self._cached(lambda: [packing.pack_depthwise(ff.dwconv.weight, self.dw16, chan_map=gmap, c_total=2 * hp)])
            self._emit("pwdw", lambda: ops.pwdw(x, pin_w, dwf_h, gated, gate=True, ln_mode=self.ln_mode, vec_t=pin_t, dw_bias=dwf_b),
                       a=x, w=pin_w, dw_w=dwf_h, out=gated, gate=True, ln_mode=self.ln_mode, vec_t=pin_t, dw_bias=dwf_b, tag="K56")
        else:
            self._gemm(x, pin_w, hid_pre, n=2 * hp, ln_mode=self.ln_mode, ln_s=pin_s, vec_t=pin_t, tag="K5")
            self._emit("dwconv", lambda: ops.dwconv3x3(hid_pre, dwf_w, gated, gate=True, bias=dwf_b), x=hid_pre, w=dwf_w, out=gated,
                       gate=True, bias=dwf_b, tag="K6")
        self._gemm(gated, pout_w, x, n=c, res=x, vec_t=pout_t, tag="K7")

    def _xblock(self, blk, x: Tensor) -> None:
        """One X-TransformerBlock on the NHWC view x (updated in place).  prompt_xrestormer.py:255-260."""
        dt = self.dtype
        B, h, w, c = x.shape
        at, sa = blk.channel_attn, blk.spatial_attn
        heads = at.num_heads
        n1, n3 = blk.norm1.body, blk.norm3.body
        beta = lambda n: getattr(n, "bias", None)

        # ---- channel attention (MDTA) ----
        qkv_w, qkv_s, qkv_t = self._cached(lambda: list(packing.pack_pointwise(at.qkv.weight, dt, gamma=n1.weight, beta=beta(n1), bias=at.qkv.bias)))
        dwq_w, dwq_b = self._cached(lambda: [packing.pack_depthwise(at.qkv_dwconv.weight, dt),
                                             None if at.qkv_dwconv.bias is None else at.qkv_dwconv.bias.detach().float().contiguous()])
        temp, wo, wo_b = self._cached(lambda: [at.temperature.detach().float().reshape(-1).contiguous(),
                                               at.project_out.weight.detach().float().reshape(c, c).contiguous(),
                                               None if at.project_out.bias is None else at.project_out.bias.detach().float().contiguous()])
        qkv_pre = self._scratch(self.S1, h, w, 3 * c)
        qkv = self._scratch(self.S2, h, w, 3 * c)
        wfold = self.wfold[c]
        splits = ops.mdta_splits(B, h * w, c)
        if self.fuse and ops.pwdw_supported(c, 3 * c, False):
            (dwq_h,) = self._cached(lambda: [packing.pack_depthwise(at.qkv_dwconv.weight, self.dw16)])
            self._emit("pwdw", lambda: ops.pwdw(x, qkv_w, dwq_h, qkv, gate=False, ln_mode=self.ln_mode, vec_t=qkv_t, dw_bias=dwq_b),
                       a=x, w=qkv_w, dw_w=dwq_h, out=qkv, gate=False, ln_mode=self.ln_mode, vec_t=qkv_t, dw_bias=dwq_b, tag="K12")
        else:
            self._gemm(x, qkv_w, qkv_pre, n=3 * c, ln_mode=self.ln_mode, ln_s=qkv_s, vec_t=qkv_t, tag="K1")
            self._emit("dwconv", lambda: ops.dwconv3x3(qkv_pre, dwq_w, qkv, gate=False, bias=dwq_b), x=qkv_pre, w=dwq_w, out=qkv,
                       gate=False, bias=dwq_b, tag="K2")
        gram_fin = ops.mdta(qkv, heads, self.ws, temp, wo, wfold, splits) if self.cuda else (None, None)
        self._emit("mdta_gram", (lambda: gram_fin[0]), qkv=qkv, heads=heads, ws=self.ws, splits=splits, tag="K3a")
        self._emit("mdta_finalize", (lambda: gram_fin[1]), qkv=qkv, heads=heads, ws=self.ws, splits=splits, temperature=temp, wo=wo,
                   wfold=wfold, tag="K3b")
        self._gemm(qkv[..., 2 * c:], wfold, x, n=c, res=x, vec_t=wo_b, w_batched=True, tag="K4")
        # ---- channel FFN ----
        self._ffn(blk.channel_ffn, blk.norm2, x)
        # ---- spatial attention (OCAB) ----
        inner, sh = sa.inner_dim, sa.num_spatial_heads
        sq_w, sq_s, sq_t = self._cached(lambda: list(packing.pack_pointwise(sa.qkv.weight, dt, gamma=n3.weight, beta=beta(n3), bias=sa.qkv.bias)))
        so_w, _, so_t = self._cached(lambda: list(packing.pack_pointwise(sa.project_out.weight, dt, bias=sa.project_out.bias)))
        rel_h, rel_w = self._cached(lambda: [sa.rel_pos_emb.rel_height.detach().float().contiguous(),
                                             sa.rel_pos_emb.rel_width.detach().float().contiguous()])
        sqkv = self._scratch(self.S1, h, w, 3 * inner)
        satt = self._scratch(self.S2, h, w, inner)
        self._gemm(x, sq_w, sqkv, n=3 * inner, ln_mode=self.ln_mode, ln_s=sq_s, vec_t=sq_t, tag="S1")
        self._emit("ocab", lambda: ops.ocab(sqkv, rel_h, rel_w, satt, heads=sh, dim_head=sa.dim_head, ws=sa.window_size, ows=sa.overlap_win_size),
                   qkv=sqkv, rel_h=rel_h, rel_w=rel_w, out=satt, heads=sh, tag="S2")
        self._gemm(satt, so_w, x, n=c, res=x, vec_t=so_t, tag="S3")
        # ---- spatial FFN ----
        self._ffn(blk.spatial_ffn, blk.norm4, x)

    def _xprompt(self, pg, cat: Tensor, lin: int, out: Tensor) -> None:
        """PromptBlock (prompt_xrestormer.py:343-359); cat[..., :lin] already holds the incoming feature."""
        B, h, w, _ = cat.shape
        d = pg.conv3x3.in_channels
        x = cat[..., :lin]
        prm, lw, lb = self._cached(lambda: [packing.pack_prompt(pg.prompt_param), pg.linear_layer.weight.detach().float().contiguous(),
                                            pg.linear_layer.bias.detach().float().contiguous()])
        (cw, cw2) = self._cached(lambda: [packing.pack_conv3x3(pg.conv3x3.weight, self.dtype), packing.pack_conv3x3(pg.conv.weight, self.dtype)])
        tmp = self._scratch(self.S2, h, w, d)
        self._emit("prompt", lambda: ops.prompt_gen(x, prm, lw, lb, tmp, self.ws, align_corners=True), x=x, prompt=prm, lin_w=lw, lin_b=lb,
                   out=tmp, ws=self.ws, align_corners=True, tag="K10")
        self._gemm(tmp, cw, cat[..., lin:], n=d, taps=9, tag="prompt_conv")
        self._xblock(pg.attn, cat)
        self._gemm(cat, cw2, out, n=lin, taps=9, tag="prompt_out")

    def kernels_per_forward(self) -> int:
        per = {"mdta_finalize": 2, "prompt": 2}
        return sum(getattr(r.get("launch"), "kernels", per.get(r["kind"], 1)) for r in self.ops)


def x_op_cost(rec: dict):
    """Algorithmic (bytes, FLOPs) of one launch; adds the OCAB core to engine.op_cost."""
    from .engine import op_cost
    if rec["kind"] == "ocab":
        q, out = rec["qkv"], rec["out"]
        pix = q.shape[0] * q.shape[1] * q.shape[2]
        inner = out.shape[3]
        return q.numel() * 2 + out.numel() * 2, 2.0 * pix * inner * 144 * 2 + 2.0 * pix * inner * 15 / 4
    return op_cost(rec)

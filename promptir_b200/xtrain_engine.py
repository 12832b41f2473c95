"""Training program of PromptXRestormer (net/prompt_xrestormer.py): forward that keeps activations + hand-derived backward.

Same machinery as `train_engine.TrainEngine` (the MDTA / GDFN halves of a block, resampling convs, prompt generation, weight
gradients and the flat gradient buffer are shared); this file adds the X block's spatial attention and the PromptBlock wiring:

  spatial attention forward   ln_fwd -> gemm (C -> 3*inner) -> pir_ocab -> gemm (inner -> C) + residual        [keeps xhat, rstd, sqkv, satt]
  spatial attention backward  dsatt = g . Wo^T | wgrad(g, satt) | pir_ocab_bwd (dq, dk, dv, d rel_h, d rel_w) | dxhat = dsqkv . Wqkv^T |
                              wgrad(dsqkv, xhat) -> dWqkv, dgamma3, dbeta3 | g += ln_bwd
  PromptBlock (prompt_xrestormer.py:343-359): prompt_gen(align_corners=True) -> conv3x3 -> [x | prompt] -> X block -> conv3x3
"""
from __future__ import annotations

from typing import Callable, Dict, List  # noqa: F401

import torch

from . import ops, packing
from ._lib import LN_BIASFREE, LN_WITHBIAS, OUT_FINAL_NCHW32
from .train_engine import TrainEngine

Tensor = torch.Tensor


class XTrainEngine(TrainEngine):
    def _build(self) -> None:
        m, B, H, W, dt = self.m, self.B, self.H, self.W, self.dtype
        if H % 64 or W % 64:
            raise ValueError("PromptXRestormer needs height and width that are multiples of 64")
        dim = m.patch_embed.proj.out_channels
        size = [(H >> l, W >> l) for l in range(4)]
        self.ln_mode = LN_BIASFREE if m.layernorm_type == "BiasFree" else LN_WITHBIAS
        self._bwd_stack: List[Callable[[], None]] = []
        self._wg_need = 0
        self._dry = False
        self.wg_ws = None
        self.saved_bytes = 0

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

        stage_levels = [(m.encoder_level1, 0), (m.encoder_level2, 1), (m.encoder_level3, 2), (m.latent, 3), (m.prompt3.attn, 3),
                        (m.decoder_level3, 2), (m.prompt2.attn, 2), (m.decoder_level2, 1), (m.prompt1.attn, 1), (m.decoder_level1, 0),
                        (m.refinement, 0)]
        ta = tb = td = 0
        widths = set()
        for mod, lvl in stage_levels:
            for blk in blocks_of(mod):
                c = blk.channel_attn.qkv.in_channels
                hp = packing.round_up(blk.channel_ffn.project_out.in_channels, 8)
                inner = blk.spatial_attn.inner_dim
                n = B * size[lvl][0] * size[lvl][1]
                ta, tb, td = max(ta, n * max(hp, inner)), max(tb, n * max(3 * c, 2 * hp, 3 * inner)), max(td, n * c)
                widths.add(c)
        for pg, lvl in ((m.prompt3, 3), (m.prompt2, 2), (m.prompt1, 1)):
            tb = max(tb, B * size[lvl][0] * size[lvl][1] * pg.conv3x3.in_channels)
        for up, lvl in ((m.up4_3, 3), (m.up3_2, 2), (m.up2_1, 1)):
            tb = max(tb, B * size[lvl][0] * size[lvl][1] * up.body[0].out_channels)
        for dn, lvl in ((m.down1_2, 0), (m.down2_3, 1), (m.down3_4, 2)):
            tb = max(tb, B * size[lvl][0] * size[lvl][1] * dn.body[0].out_channels)
        self.Ta, self.Tb, self.Tc, self.Td = (torch.zeros(n, dtype=dt, device=self.device) for n in (ta, tb, tb, td))
        self.wfold = {c: self._zeros(B, c, packing.kpad_of(c)) for c in sorted(widths)}
        self.wft = {c: self._zeros(B, c, packing.kpad_of(c)) for c in sorted(widths)}
        self.wqk = {c: self._zeros(B, 2 * c, packing.kpad_of(2 * c)) for c in sorted(widths)}

        c1, c2, c3, c4 = dim, dim * 2, dim * 4, dim * 8
        p1, p2, p3 = (pg.conv3x3.in_channels for pg in (m.prompt1, m.prompt2, m.prompt3))
        (h0, w0), (h1, w1), (h2, w2), (h3, w3) = size
        up1 = m.up2_1.body[0].out_channels // 4
        up2 = m.up3_2.body[0].out_channels // 4
        up3 = m.up4_3.body[0].out_channels // 4
        shapes = dict(cat1=(h0, w0, up1 + c1), cat2=(h1, w1, up2 + c2), cat3=(h2, w2, up3 + c3), pcat3=(h3, w3, c4 + p3),
                      pcat2=(h2, w2, c3 + p2), pcat1=(h1, w1, c2 + p1), r3=(h3, w3, c4), r2=(h2, w2, c3), r1=(h1, w1, c2),
                      dec1=(h0, w0, up1 + c1))
        for name, shp in shapes.items():
            setattr(self, name, self._zeros(B, *shp))
            if name != "cat1":
                setattr(self, "g_" + name, self._zeros(B, *shp))
        cin, cout = m.patch_embed.proj.in_channels, m.output.out_channels
        self.img_in = self._f32(B, cin, H, W)
        self.out = self._f32(B, cout, H, W)
        self.d_out = self._f32(B, cout, H, W)
        self.d_out8 = self._zeros(B, H, W, 8)
        self.img8 = self._zeros(B, H, W, 8)
        self.d_img = self._f32(B, cin, H, W) if self.input_grad else None

        enc1, enc2, enc3 = self.cat1[..., up1:], self.cat2[..., up2:], self.cat3[..., up3:]
        lat, d3, d2 = self.pcat3[..., :c4], self.pcat2[..., :c3], self.pcat1[..., :c2]
        g_enc1, g_enc2, g_enc3 = self.g_dec1[..., up1:], self.g_cat2[..., up2:], self.g_cat3[..., up3:]
        g_lat, g_d3, g_d2 = self.g_pcat3[..., :c4], self.g_pcat2[..., :c3], self.g_pcat1[..., :c2]

        # ================================ forward program (prompt_xrestormer.py:428-478) =====================================
        self.ops = self.fwd_ops = []
        pe = m.patch_embed.proj
        pe_w, pe_b = self._cached(lambda: [pe.weight.detach().float().contiguous(),
                                           None if pe.bias is None else pe.bias.detach().float().contiguous()])
        self._emit("patch_embed", lambda: ops.patch_embed(self.img_in, pe_w, pe_b, enc1), img=self.img_in, w=pe_w, bias=pe_b, out=enc1)
        self._later(lambda: self._patch_embed_bwd(pe, g_enc1))
        self._txstage(m.encoder_level1, enc1, g_enc1)
        self._tdown(m.down1_2, enc1, enc2, g_enc1, g_enc2)
        self._txstage(m.encoder_level2, enc2, g_enc2)
        self._tdown(m.down2_3, enc2, enc3, g_enc2, g_enc3)
        self._txstage(m.encoder_level3, enc3, g_enc3)
        self._tdown(m.down3_4, enc3, lat, g_enc3, g_lat)
        self._txstage(m.latent, lat, g_lat)
        self._txprompt(m.prompt3, self.pcat3, c4, self.r3, self.g_pcat3, self.g_r3)
        self._tup(m.up4_3, self.r3, self.cat3[..., :up3], self.g_r3, self.g_cat3[..., :up3])
        self._treduce(m.reduce_chan_level3, self.cat3, d3, self.g_cat3, g_d3)
        self._txstage(m.decoder_level3, d3, g_d3)
        self._txprompt(m.prompt2, self.pcat2, c3, self.r2, self.g_pcat2, self.g_r2)
        self._tup(m.up3_2, self.r2, self.cat2[..., :up2], self.g_r2, self.g_cat2[..., :up2])
        self._treduce(m.reduce_chan_level2, self.cat2, d2, self.g_cat2, g_d2)
        self._txstage(m.decoder_level2, d2, g_d2)
        self._txprompt(m.prompt1, self.pcat1, c2, self.r1, self.g_pcat1, self.g_r1)
        self._tup(m.up2_1, self.r1, self.cat1[..., :up1], self.g_r1, self.g_dec1[..., :up1])
        # decoder_level1 would overwrite the encoder half of cat1 that down1_2's weight gradient needs: first block -> dec1
        self._txstage(m.decoder_level1, self.cat1, self.g_dec1, first_out=self.dec1)
        self._txstage(m.refinement, self.dec1, self.g_dec1)
        oc = m.output
        (ow,) = self._cached(lambda: [packing.pack_conv3x3(oc.weight, dt)])
        (ob,) = self._cached(lambda: [None if oc.bias is None else oc.bias.detach().float().contiguous()])
        self._gemm(self.dec1, ow, self.out, n=oc.out_channels, taps=9, out_mode=OUT_FINAL_NCHW32, vec_t=ob, img=self.img_in, tag="output")
        self._later(lambda: self._output_bwd(oc))
        self._finish_build()

    # ---- X blocks --------------------------------------------------------------------------------------------------------
    def _txstage(self, mod, x: Tensor, g: Tensor, first_out=None) -> None:
        blocks = list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]
        for i, blk in enumerate(blocks):
            if i == 0 and first_out is not None:
                self._txblock(blk, x, g, x_out=first_out)
                x = first_out
            else:
                self._txblock(blk, x, g)

    def _txblock(self, blk, x: Tensor, g: Tensor, x_out=None) -> None:
        """prompt_xrestormer.py:255-260: channel attention, channel FFN, spatial attention, spatial FFN (each with its LayerNorm)."""
        xo = x if x_out is None else x_out
        self._t_mdta(blk.channel_attn, blk.norm1, x, g, xo)
        self._t_gdfn(blk.channel_ffn, blk.norm2, xo, g)
        self._t_ocab(blk.spatial_attn, blk.norm3, xo, g)
        self._t_gdfn(blk.spatial_ffn, blk.norm4, xo, g)

    def _t_ocab(self, sa, norm, x: Tensor, g: Tensor) -> None:
        """x += OCAB(LN(x)) in place (prompt_xrestormer.py:209-235, 258) and its backward on g."""
        dt = self.dtype
        B, h, w, c = x.shape
        inner, sh = sa.inner_dim, sa.num_spatial_heads
        n3 = norm.body
        beta3 = getattr(n3, "bias", None)
        sq_w, _, sq_t = self._cached(lambda: list(packing.pack_pointwise(sa.qkv.weight, dt, gamma=n3.weight, beta=beta3, bias=sa.qkv.bias)))
        (sq_wT,) = self._cached(lambda: [packing.pack_pointwise((sa.qkv.weight.detach().reshape(3 * inner, c) * n3.weight.detach().view(1, -1)).t(), dt)[0]])
        so_w, _, so_t = self._cached(lambda: list(packing.pack_pointwise(sa.project_out.weight, dt, bias=sa.project_out.bias)))
        (so_wT,) = self._cached(lambda: [packing.pack_pointwise(sa.project_out.weight.detach().reshape(c, inner).t(), dt)[0]])
        rel_h, rel_w = self._cached(lambda: [sa.rel_pos_emb.rel_height.detach().float().contiguous(),
                                             sa.rel_pos_emb.rel_width.detach().float().contiguous()])
        keep = lambda ch: self._zeros(B, h, w, ch)
        xh3, sqkv, satt = keep(c), keep(3 * inner), keep(inner)
        rstd3 = self._f32(B * h * w)
        self.saved_bytes += sum(t.numel() * t.element_size() for t in (xh3, sqkv, satt, rstd3))

        self._ln_fwd(x, xh3, rstd3, "LN3")
        self._gemm(xh3, sq_w, sqkv, n=3 * inner, vec_t=sq_t, tag="S1")
        self._emit("ocab", lambda: ops.ocab(sqkv, rel_h, rel_w, satt, heads=sh, dim_head=sa.dim_head, ws=sa.window_size, ows=sa.overlap_win_size),
                   qkv=sqkv, rel_h=rel_h, rel_w=rel_w, out=satt, heads=sh, tag="S2")
        self._gemm(satt, so_w, x, n=c, res=x, vec_t=so_t, tag="S3")

        def bwd():
            G = self._grad_of
            dsatt = self._scratch(self.Ta, h, w, inner)
            dsqkv = self._scratch(self.Tb, h, w, 3 * inner)
            dxh = self._scratch(self.Td, h, w, c)
            self._gemm(g, so_wT, dsatt, n=inner, tag="BS3d")
            wg = self._wgrad(g, satt, colsum=sa.project_out.bias is not None, tag="BS3w")
            self._wgrad_fin(wg, dst_w=G(sa.project_out.weight).view(c, inner), dst_bias=G(sa.project_out.bias), tag="BS3f")
            rec = dict(qkv=sqkv, dout=dsatt, rel_h=rel_h, rel_w=rel_w, dqkv=dsqkv, heads=sh, dst_rel_h=G(sa.rel_pos_emb.rel_height),
                       dst_rel_w=G(sa.rel_pos_emb.rel_width), inv_scale=1.0 / self.grad_scale, tag="BS2")
            self._wg_need = max(self._wg_need, ops.ocab_bwd_ws_floats(B, h, w, sh))
            self._emit("ocab_bwd", lambda: ops.ocab_bwd(self.wg_ws, rec), ws=self.wg_ws, **rec)
            self._gemm(dsqkv, sq_wT, dxh, n=c, tag="BS1d")
            wg = self._wgrad(dsqkv, xh3, colsum=True, tag="BS1w")
            self._wgrad_fin(wg, dst_w=G(sa.qkv.weight).view(3 * inner, c), gamma=self._raw(n3.weight), beta=self._raw(beta3),
                            w=self._raw(sa.qkv.weight), dst_gamma=G(n3.weight), dst_beta=G(beta3), dst_bias=G(sa.qkv.bias), tag="BS1f")
            self._ln_bwd(dxh, xh3, rstd3, g, "BLN3")
        self._later(bwd)

    def _txprompt(self, pg, cat: Tensor, lin: int, out: Tensor, g_cat: Tensor, g_out: Tensor) -> None:
        """PromptBlock (prompt_xrestormer.py:343-359).  cat[..., :lin] already holds the incoming feature; out is the block's result."""
        x, g_x = cat[..., :lin], g_cat[..., :lin]
        self._tprompt(pg, x, cat[..., lin:], g_x, g_cat[..., lin:], align_corners=True)
        self._txblock(pg.attn, cat, g_cat)
        self._tconv3(pg.conv, cat, out, g_cat, g_out, tag="pconv")

    def kernels_per_step(self) -> int:
        per = {"mdta_finalize": 2, "prompt": 2, "wgrad_fin": 2, "dw_wgrad": 2, "mdta_bwd": 6, "prompt_bwd": 4, "ocab_bwd": 3}
        return sum(getattr(r.get("launch"), "kernels", per.get(r["kind"], 1)) for r in self.ops)

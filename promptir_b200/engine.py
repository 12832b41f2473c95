"""Execution engine of the PromptIR forward on one B200: buffer plan, packed-weight cache and the launch program.

One Engine is specialised to (batch, height, width, dtype).  It owns
  * the NHWC 16-bit activation arena.  torch.cat (net/model.py:341,347,353,359,365,370) never happens:
    every concatenation target is allocated once and its producers write straight into channel slices of
    it (all kernels take a pixel pitch).  The residual stream of a stage is updated in place.
  * derived weight caches (packing.py) -- rebuilt in place whenever a parameter's version counter changes,
  * the program: an ordered list of prepared C-ABI launches (ops.py), replayed eagerly or as a CUDA graph.

Per TransformerBlock (net/model.py:192-196) the program is five kernels (+ two tiny ones) where the x tile fits
shared memory (C <= 192), seven otherwise:
  K12 pwdw  LN1 + 1x1 (C -> 3C) + dw3x3             K56 pwdw  LN2 + 1x1 (C -> 2*hp) + dw3x3 + GELU gate
      (else K1 gemm, K2 dwconv)                         (else K5 gemm, K6 dwconv)
  K3 mdta_gram (+ finalize: softmax, fold into Wo)  K7 gemm   hp -> C, + residual (in place)
  K4 gemm   v . Wfold[b] -> C, + residual (in place)

The op records in `self.ops` are plain dicts (kind + tensor views + scalars).  On a CUDA device each record also
carries a prepared launch; tests interpret the same records with a torch emulator on CPU to check the wiring
and the packing without a GPU.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch

from . import ops, packing
from .repack import Packer
from ._lib import (LN_BIASFREE, LN_NONE, LN_WITHBIAS, OUT_FINAL_NCHW32, OUT_NHWC16, OUT_SHUFFLE16, OUT_UNSHUFFLE16)

Tensor = torch.Tensor


class Engine:
    def __init__(self, module, batch: int, height: int, width: int, device, dtype: torch.dtype = torch.bfloat16,
                 img_in: Optional[Tensor] = None, out: Optional[Tensor] = None, fuse: Optional[bool] = None):
        if height % 8 or width % 8:
            raise ValueError("height and width must be multiples of 8")
        self.m = module
        self.B, self.H, self.W = batch, height, width
        self.device = torch.device(device)
        self.dtype = dtype
        self.cuda = self.device.type == "cuda"
        # depthwise taps of the fused kernels are IEEE half whatever the storage type (fp32 only for the CPU wiring tests)
        self.dw16 = torch.float32 if dtype == torch.float32 else torch.float16
        # fuse=False / PROMPTIR_B200_FUSE=0: never use the fused pir_pwdw kernels (A/B timing, diagnostics.range_report)
        self.fuse = (os.environ.get("PROMPTIR_B200_FUSE", "1") != "0") if fuse is None else fuse
        # q|k and v of MDTA (model.py:121) as two dense tensors written by the fused qkv kernel: the Gram streams 2C-channel rows and
        # the attn.v GEMM C-channel rows instead of slices of 3C-channel rows.  0: one [.., 3C] tensor (A/B timing)
        self.split_qkv = os.environ.get("PROMPTIR_B200_SPLITQKV", "1") != "0"
        if self.cuda:
            from . import _lib
            _lib.check(_lib.load().pir_check_device(), "pir_check_device")
        self.ops: List[dict] = []
        self._packers: List[Callable[[], None]] = []
        # derived weight caches: requests recorded while the program is built, filled / refreshed by ONE pir_repack launch
        self.pk = Packer(self.device, verify=os.environ.get("PROMPTIR_B200_PACK_VERIFY") == "1")
        self._graph = None
        self._param_version = -1
        self._io = (img_in, out)
        self._build()
        if self.cuda:
            self.pk.run()
        self._ptr_fingerprint = self._pointer_fingerprint()
        self._param_version = self._current_version()

    # ------------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------------
    def _zeros(self, *shape, dtype=None) -> Tensor:
        return torch.zeros(*shape, dtype=dtype or self.dtype, device=self.device)

    def _cached(self, make: Callable[[], List[Optional[Tensor]]]) -> List[Optional[Tensor]]:
        """Materialise packed tensors now and register a refresher that rewrites them in place."""
        cur = make()

        def refresh(cur=cur, make=make):
            for dst, src in zip(cur, make()):
                if dst is not None:
                    dst.copy_(src)
        self._packers.append(refresh)
        return cur

    def _emit(self, kind: str, launch_fn: Optional[Callable] = None, **args) -> None:
        rec = {"kind": kind, **args}
        rec["launch"] = launch_fn() if (self.cuda and launch_fn is not None) else None
        self.ops.append(rec)

    def _gemm(self, a, w, out, *, n, taps=1, out_mode=OUT_NHWC16, res=None, ln_mode=LN_NONE, ln_s=None, vec_t=None,
              img=None, w_batched=False, tag=""):
        self._emit("gemm", lambda: ops.gemm(a, w, out, n=n, taps=taps, out_mode=out_mode, res=res, ln_mode=ln_mode,
                                            ln_s=ln_s, vec_t=vec_t, img=img, w_batched=w_batched),
                   a=a, w=w, out=out, n=n, taps=taps, out_mode=out_mode, res=res, ln_mode=ln_mode, ln_s=ln_s,
                   vec_t=vec_t, img=img, w_batched=w_batched, tag=tag)

    def _scratch(self, flat: Tensor, h: int, w: int, c: int) -> Tensor:
        n = self.B * h * w * c
        assert n <= flat.numel(), "scratch arena too small"
        return flat[:n].view(self.B, h, w, c)

    # ------------------------------------------------------------------------------------------------
    # program construction
    # ------------------------------------------------------------------------------------------------
    def _build(self) -> None:
        m, B, H, W, dt = self.m, self.B, self.H, self.W, self.dtype
        dim = m.patch_embed.proj.out_channels
        size = [(H >> l, W >> l) for l in range(4)]
        self.ln_mode = LN_BIASFREE if m.layernorm_type == "BiasFree" else LN_WITHBIAS

        # ---- stage table: (module, level) in execution order is wired below; first size the arenas ----
        def blocks_of(mod):
            return list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]

        stage_levels = [(m.encoder_level1, 0), (m.encoder_level2, 1), (m.encoder_level3, 2), (m.latent, 3),
                        (m.noise_level3, 3), (m.decoder_level3, 2), (m.noise_level2, 2), (m.decoder_level2, 1),
                        (m.noise_level1, 1), (m.decoder_level1, 0), (m.refinement, 0)]
        s1 = s2 = ws_f = 0
        widths = set()
        for mod, lvl in stage_levels:
            for blk in blocks_of(mod):
                c = blk.attn.qkv.in_channels
                hp = packing.round_up(blk.ffn.project_out.in_channels, 8)
                hw = size[lvl][0] * size[lvl][1]
                s1 = max(s1, B * hw * max(3 * c, 2 * hp))
                s2 = max(s2, B * hw * max(3 * c, hp))
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
        # folded attention weights: pad columns must stay zero -> one zero-initialised buffer per width
        self.wfold: Dict[int, Tensor] = {c: self._zeros(B, c, packing.kpad_of(c)) for c in sorted(widths)}

        c1, c2, c3, c4 = dim, dim * 2, dim * 4, dim * 8
        p1, p2, p3 = (pg.conv3x3.in_channels for pg in (m.prompt1, m.prompt2, m.prompt3))
        (h0, w0), (h1, w1), (h2, w2), (h3, w3) = size
        up1 = m.up2_1.body[0].out_channels // 4          # channels after PixelShuffle
        up2 = m.up3_2.body[0].out_channels // 4
        up3 = m.up4_3.body[0].out_channels // 4
        self.cat1 = self._zeros(B, h0, w0, up1 + c1)      # [up2_1 | encoder_level1]      model.py:370
        self.cat2 = self._zeros(B, h1, w1, up2 + c2)      # [up3_2 | encoder_level2]      model.py:359
        self.cat3 = self._zeros(B, h2, w2, up3 + c3)      # [up4_3 | encoder_level3]      model.py:347
        self.catn3 = self._zeros(B, h3, w3, c4 + p3)      # [latent | prompt3]            model.py:341
        self.catn2 = self._zeros(B, h2, w2, c3 + p2)      # [decoder_level3 | prompt2]    model.py:353
        self.catn1 = self._zeros(B, h1, w1, c2 + p1)      # [decoder_level2 | prompt1]    model.py:365
        self.r3 = self._zeros(B, h3, w3, m.reduce_noise_level3.out_channels)
        self.r2 = self._zeros(B, h2, w2, m.reduce_noise_level2.out_channels)
        self.r1 = self._zeros(B, h1, w1, m.reduce_noise_level1.out_channels)
        self.img_in = self._io[0] if self._io[0] is not None else torch.zeros(B, m.patch_embed.proj.in_channels, H, W, dtype=torch.float32,
                                                                               device=self.device)
        self.out = self._io[1] if self._io[1] is not None else torch.zeros(B, m.output.out_channels, H, W, dtype=torch.float32,
                                                                            device=self.device)
        assert self.img_in.is_contiguous() and self.out.is_contiguous() and self.img_in.shape[0] == B and self.out.shape[0] == B

        enc1 = self.cat1[..., up1:]
        enc2 = self.cat2[..., up2:]
        enc3 = self.cat3[..., up3:]
        lat = self.catn3[..., :c4]
        d3 = self.catn2[..., :c3]
        d2 = self.catn1[..., :c2]

        # ---- encoder -----------------------------------------------------------------------------------
        pe = m.patch_embed.proj
        pe_w, pe_b = self.pk.f32(pe.weight), self.pk.f32(pe.bias)
        self._emit("patch_embed", lambda: ops.patch_embed(self.img_in, pe_w, pe_b, enc1), img=self.img_in, w=pe_w, bias=pe_b,
                   out=enc1)
        self._stage(m.encoder_level1, enc1)
        self._down(m.down1_2, enc1, enc2)
        self._stage(m.encoder_level2, enc2)
        self._down(m.down2_3, enc2, enc3)
        self._stage(m.encoder_level3, enc3)
        self._down(m.down3_4, enc3, lat)
        self._stage(m.latent, lat)
        # ---- decoder with prompts ----------------------------------------------------------------------
        self._prompt(m.prompt3, lat, self.catn3[..., c4:])
        self._stage(m.noise_level3, self.catn3)
        self._reduce(m.reduce_noise_level3, self.catn3, self.r3)
        self._up(m.up4_3, self.r3, self.cat3[..., :up3])
        self._reduce(m.reduce_chan_level3, self.cat3, d3)
        self._stage(m.decoder_level3, d3)
        self._prompt(m.prompt2, d3, self.catn2[..., c3:])
        self._stage(m.noise_level2, self.catn2)
        self._reduce(m.reduce_noise_level2, self.catn2, self.r2)
        self._up(m.up3_2, self.r2, self.cat2[..., :up2])
        self._reduce(m.reduce_chan_level2, self.cat2, d2)
        self._stage(m.decoder_level2, d2)
        self._prompt(m.prompt1, d2, self.catn1[..., c2:])
        self._stage(m.noise_level1, self.catn1)
        self._reduce(m.reduce_noise_level1, self.catn1, self.r1)
        self._up(m.up2_1, self.r1, self.cat1[..., :up1])
        self._stage(m.decoder_level1, self.cat1)
        self._stage(m.refinement, self.cat1)
        oc = m.output
        ow, ob = self.pk.conv3x3(oc.weight, dt), self.pk.f32(oc.bias)
        self._gemm(self.cat1, ow, self.out, n=oc.out_channels, taps=9, out_mode=OUT_FINAL_NCHW32, vec_t=ob, img=self.img_in,
                   tag="output")
        self.launches = [r["launch"] for r in self.ops]

    def _stage(self, mod, x: Tensor) -> None:
        for blk in (list(mod) if isinstance(mod, torch.nn.Sequential) else [mod]):
            self._block(blk, x)

    def _block(self, blk, x: Tensor) -> None:
        """One TransformerBlock on the NHWC view x (updated in place).  net/model.py:192-196."""
        dt = self.dtype
        B, h, w, c = x.shape
        heads = blk.attn.num_heads
        hid = blk.ffn.project_out.in_channels
        hp, gmap = packing.gdfn_maps(hid, self.device)
        n1, n2 = blk.norm1.body, blk.norm2.body
        at, ff = blk.attn, blk.ffn
        beta = lambda n: getattr(n, "bias", None)

        pk = self.pk
        qkv_w, qkv_s, qkv_t = pk.pointwise(at.qkv.weight, dt, gamma=n1.weight, beta=beta(n1), bias=at.qkv.bias)
        dwq_b = pk.f32(at.qkv_dwconv.bias)
        temp, wo, wo_b = pk.f32(at.temperature, (-1,)), pk.f32(at.project_out.weight, (c, c)), pk.f32(at.project_out.bias)
        pin_w, pin_s, pin_t = pk.pointwise(ff.project_in.weight, dt, gamma=n2.weight, beta=beta(n2), bias=ff.project_in.bias,
                                           rows=(hid, hp), n_total=2 * hp)
        dwf_b = pk.vec(ff.dwconv.bias, (hid, hp), 2 * hp)
        pout_w, _, pout_t = pk.pointwise(ff.project_out.weight, dt, bias=ff.project_out.bias, k_total=hp)

        qkv_pre = self._scratch(self.S1, h, w, 3 * c)
        qkv = self._scratch(self.S2, h, w, 3 * c)
        hid_pre = self._scratch(self.S1, h, w, 2 * hp)
        gated = self._scratch(self.S2, h, w, hp)
        wfold = self.wfold[c]
        splits = ops.mdta_splits(B, h * w, c)
        # fused LN -> 1x1 -> depthwise 3x3 kernels where the x tile fits shared memory (levels 1-3); else two kernels
        fuse_qkv = self.fuse and ops.pwdw_supported(c, 3 * c, False)
        fuse_ffn = self.fuse and ops.pwdw_supported(c, hp, True)

        split = fuse_qkv and self.split_qkv and ops.pwdw_split_supported(c, 3 * c)
        qk, v = qkv[..., :2 * c], qkv[..., 2 * c:]
        if split:
            n = B * h * w
            qk = self.S2[:n * 2 * c].view(B, h, w, 2 * c)
            v = self.S2[n * 2 * c:n * 3 * c].view(B, h, w, c)
        if fuse_qkv:
            dwq_h = pk.depthwise(at.qkv_dwconv.weight, self.dw16)
            o1, o2 = (qk, v) if split else (qkv, None)
            self._emit("pwdw", lambda: ops.pwdw(x, qkv_w, dwq_h, o1, gate=False, ln_mode=self.ln_mode, vec_t=qkv_t, dw_bias=dwq_b, out2=o2),
                       a=x, w=qkv_w, dw_w=dwq_h, out=o1, out2=o2, gate=False, ln_mode=self.ln_mode, vec_t=qkv_t, dw_bias=dwq_b, tag="K12")
        else:
            dwq_w = pk.depthwise(at.qkv_dwconv.weight, dt)
            self._gemm(x, qkv_w, qkv_pre, n=3 * c, ln_mode=self.ln_mode, ln_s=qkv_s, vec_t=qkv_t, tag="K1")
            self._emit("dwconv", lambda: ops.dwconv3x3(qkv_pre, dwq_w, qkv, gate=False, bias=dwq_b), x=qkv_pre, w=dwq_w, out=qkv,
                       gate=False, bias=dwq_b, tag="K2")
        gram_fin = ops.mdta(qk, heads, self.ws, temp, wo, wfold, splits, qk_only=True) if self.cuda else (None, None)
        self._emit("mdta_gram", (lambda: gram_fin[0]), qk=qk, heads=heads, ws=self.ws, splits=splits, tag="K3a")
        self._emit("mdta_finalize", (lambda: gram_fin[1]), qk=qk, heads=heads, ws=self.ws, splits=splits, temperature=temp,
                   wo=wo, wfold=wfold, tag="K3b")
        self._gemm(v, wfold, x, n=c, res=x, vec_t=wo_b, w_batched=True, tag="K4")
        if fuse_ffn:
            dwf_h = pk.depthwise(ff.dwconv.weight, self.dw16, split=(hid, hp), c_total=2 * hp)
            self._emit("pwdw", lambda: ops.pwdw(x, pin_w, dwf_h, gated, gate=True, ln_mode=self.ln_mode, vec_t=pin_t, dw_bias=dwf_b),
                       a=x, w=pin_w, dw_w=dwf_h, out=gated, gate=True, ln_mode=self.ln_mode, vec_t=pin_t, dw_bias=dwf_b, tag="K56")
        else:
            dwf_w = pk.depthwise(ff.dwconv.weight, dt, split=(hid, hp), c_total=2 * hp)
            self._gemm(x, pin_w, hid_pre, n=2 * hp, ln_mode=self.ln_mode, ln_s=pin_s, vec_t=pin_t, tag="K5")
            self._emit("dwconv", lambda: ops.dwconv3x3(hid_pre, dwf_w, gated, gate=True, bias=dwf_b), x=hid_pre, w=dwf_w, out=gated,
                       gate=True, bias=dwf_b, tag="K6")
        self._gemm(gated, pout_w, x, n=c, res=x, vec_t=pout_t, tag="K7")

    def _down(self, mod, x: Tensor, out: Tensor) -> None:
        conv = mod.body[0]                                  # model.py:164-165
        w = self.pk.conv3x3(conv.weight, self.dtype)
        self._gemm(x, w, out, n=conv.out_channels, taps=9, out_mode=OUT_UNSHUFFLE16, tag="down")

    def _up(self, mod, x: Tensor, out: Tensor) -> None:
        conv = mod.body[0]                                  # model.py:174-175
        w = self.pk.conv3x3(conv.weight, self.dtype)
        self._gemm(x, w, out, n=conv.out_channels, taps=9, out_mode=OUT_SHUFFLE16, tag="up")

    def _reduce(self, conv, x: Tensor, out: Tensor) -> None:
        w, _, t = self.pk.pointwise(conv.weight, self.dtype, bias=conv.bias)
        self._gemm(x, w, out, n=conv.out_channels, vec_t=t, tag="reduce")

    def _prompt(self, pg, x: Tensor, out: Tensor) -> None:
        """PromptGenBlock (model.py:226-235): fused pool/linear/softmax/mix/bilinear, then the 3x3 conv."""
        B, h, w, _ = x.shape
        d = pg.conv3x3.in_channels
        prm, lw, lb = self.pk.prompt(pg.prompt_param), self.pk.f32(pg.linear_layer.weight), self.pk.f32(pg.linear_layer.bias)
        cw = self.pk.conv3x3(pg.conv3x3.weight, self.dtype)
        tmp = self._scratch(self.S2, h, w, d)
        self._emit("prompt", lambda: ops.prompt_gen(x, prm, lw, lb, tmp, self.ws), x=x, prompt=prm, lin_w=lw, lin_b=lb, out=tmp,
                   ws=self.ws, tag="K10")
        self._gemm(tmp, cw, out, n=d, taps=9, tag="prompt_conv")

    # ------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------
    def _current_version(self) -> int:
        return sum(p._version for p in self.m.parameters())

    def params_moved(self) -> bool:
        """True when a parameter's storage was replaced since the engine was built (`p.data = ...`, load_state_dict(assign=True),
        EMA swaps): the launch descriptors hold raw pointers to the fp32 parameters, so the owner must build a new engine."""
        return self.cuda and self._ptr_fingerprint != self._pointer_fingerprint()

    def _pointer_fingerprint(self) -> int:
        h = 0
        for p in self.m.parameters():                   # the module's CURRENT parameter objects (assign=True swaps the objects too)
            h = (h * 1000003 + p.data_ptr()) & 0xFFFFFFFFFFFF
        return h

    def refresh_weights(self) -> None:
        """Rebuild the packed caches in place (keeps pointers, so a captured graph stays valid): one pir_repack launch."""
        self.pk.run()
        with torch.no_grad():
            for fn in self._packers:            # torch packers of subclasses that still use _cached
                fn()
        self._param_version = self._current_version()

    def launch_all(self, stream: int) -> None:
        for fn in self.launches:
            fn(stream)

    def run(self, img: Tensor, use_graph: bool = True) -> Tensor:
        """img: fp32 NCHW contiguous on this engine's device -> restored fp32 NCHW (a fresh tensor)."""
        if not self.cuda:
            raise RuntimeError("promptir_b200.Engine.run needs a CUDA (sm_100a) device; there is no CPU path")
        if tuple(img.shape) != tuple(self.img_in.shape):
            raise ValueError(f"engine built for {tuple(self.img_in.shape)}, got {tuple(img.shape)}")
        if self._current_version() != self._param_version:
            self.refresh_weights()
        self.img_in.copy_(img)
        self.replay(use_graph)
        return self.out.clone()

    def replay(self, use_graph: bool = True) -> None:
        """Run the program on whatever is in self.img_in; result lands in self.out (no copies)."""
        if use_graph:
            if self._graph is None:
                self._capture()
            self._graph.replay()
        else:
            self.launch_all(torch.cuda.current_stream(self.device).cuda_stream)

    def _capture(self) -> None:
        # one eager pass first: sets kernel attributes / loads modules outside of capture
        self.launch_all(torch.cuda.current_stream(self.device).cuda_stream)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch_all(torch.cuda.current_stream(self.device).cuda_stream)
        self._graph = g

    # bookkeeping for bench.py -------------------------------------------------------------------------
    def kernels_per_forward(self) -> int:
        per = {"mdta_finalize": 2, "prompt": 2}
        return sum(getattr(r.get("launch"), "kernels", per.get(r["kind"], 1)) for r in self.ops)


class SplitEngine:
    """Two half-batch Engines replayed as two parallel branches of ONE CUDA graph.

    Images are independent, so the two halves have no dependency on each other: while one branch sits in a latency-bound
    step (the tiny MDTA finalize kernels, the 32x32-level blocks, kernel tails) the other branch's persistent kernels take
    the idle SMs.  Same interface as Engine (img_in / out are the full-batch tensors; the halves work on views)."""

    def __init__(self, module, batch: int, height: int, width: int, device, dtype: torch.dtype = torch.bfloat16):
        assert batch >= 2
        self.m, self.B, self.H, self.W = module, batch, height, width
        self.device, self.dtype = torch.device(device), dtype
        self.cuda = self.device.type == "cuda"
        cin, cout = module.patch_embed.proj.in_channels, module.output.out_channels
        self.img_in = torch.zeros(batch, cin, height, width, dtype=torch.float32, device=self.device)
        self.out = torch.zeros(batch, cout, height, width, dtype=torch.float32, device=self.device)
        h = batch // 2
        self.parts = [Engine(module, h, height, width, device, dtype, self.img_in[:h], self.out[:h]),
                      Engine(module, batch - h, height, width, device, dtype, self.img_in[h:], self.out[h:])]
        self.ops = self.parts[0].ops + self.parts[1].ops
        self.launches = self.parts[0].launches + self.parts[1].launches
        self._graph = None
        self._side = torch.cuda.Stream(self.device) if self.cuda else None

    def refresh_weights(self) -> None:
        for p in self.parts:
            p.refresh_weights()

    def launch_all(self, stream: int) -> None:
        for p in self.parts:
            p.launch_all(stream)

    def kernels_per_forward(self) -> int:
        return sum(p.kernels_per_forward() for p in self.parts)

    def run(self, img: Tensor, use_graph: bool = True) -> Tensor:
        if not self.cuda:
            raise RuntimeError("promptir_b200.SplitEngine.run needs a CUDA (sm_100a) device; there is no CPU path")
        if tuple(img.shape) != tuple(self.img_in.shape):
            raise ValueError(f"engine built for {tuple(self.img_in.shape)}, got {tuple(img.shape)}")
        for p in self.parts:
            if p._current_version() != p._param_version:
                p.refresh_weights()
        self.img_in.copy_(img)
        self.replay(use_graph)
        return self.out.clone()

    def replay(self, use_graph: bool = True) -> None:
        if not use_graph:
            self.launch_all(torch.cuda.current_stream(self.device).cuda_stream)
            return
        if self._graph is None:
            self.launch_all(torch.cuda.current_stream(self.device).cuda_stream)     # warm: kernel attributes outside capture
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                main = torch.cuda.current_stream(self.device)
                self._side.wait_stream(main)                                         # fork
                self.parts[0].launch_all(main.cuda_stream)
                with torch.cuda.stream(self._side):
                    self.parts[1].launch_all(self._side.cuda_stream)
                main.wait_stream(self._side)                                         # join
            self._graph = g
        self._graph.replay()


# ----------------------------------------------------------------------------------------------------
# algorithmic cost of an op record (for roofline accounting; DESIGN.md states the per-unit figures)
# ----------------------------------------------------------------------------------------------------
def _numel(t) -> int:
    return 0 if t is None else int(t.numel())


def _qk_dims(rec: dict):
    """(B, H, W, C) of an MDTA record: it carries either the dense / sliced q|k tensor [.., 2C] or the whole qkv tensor [.., 3C]."""
    if rec.get("qk") is not None:
        B, H, W, c2 = rec["qk"].shape
        return B, H, W, c2 // 2
    B, H, W, c3 = rec["qkv"].shape
    return B, H, W, c3 // 3


def op_cost(rec: dict):
    """-> (algorithmic HBM bytes, FLOPs) of one launch: every operand read once, every result written once."""
    kind = rec["kind"]
    if kind == "gemm":
        a, out = rec["a"], rec["out"]
        B, H, W, K = a.shape
        n, taps = rec["n"], rec["taps"]
        wbytes = (B if rec["w_batched"] else 1) * n * taps * K * 2
        by = a.numel() * a.element_size() + _numel(out) * out.element_size() + wbytes
        if rec["res"] is not None:
            by += _numel(rec["res"]) * rec["res"].element_size()
        if rec["img"] is not None:
            by += _numel(rec["img"]) * 4
        return by, 2.0 * B * H * W * K * taps * n
    if kind == "pwdw":
        a, out = rec["a"], rec["out"]
        npre = rec["w"].shape[0]
        pix = a.shape[0] * a.shape[1] * a.shape[2]
        return (a.numel() * 2 + (out.numel() + _numel(rec.get("out2"))) * 2 + rec["w"].numel() * 2 + rec["dw_w"].numel() * 2,
                2.0 * pix * a.shape[3] * npre + 18.0 * pix * npre)
    if kind == "dwconv":
        x, out = rec["x"], rec["out"]
        return x.numel() * 2 + out.numel() * 2 + rec["w"].numel() * 2, 18.0 * x.numel()
    if kind == "mdta_gram":
        B, H, W, c = _qk_dims(rec)
        return B * H * W * 2 * c * 2, 2.0 * B * H * W * c * (c // rec["heads"])
    if kind == "mdta_finalize":
        B, _, _, c = _qk_dims(rec)
        return B * c * c * (4 * rec["splits"] + 2) + c * c * 4, 2.0 * B * c * c * (c // rec["heads"])
    if kind == "prompt":
        x, out = rec["x"], rec["out"]
        return x.numel() * 2 + out.numel() * 2 + rec["prompt"].numel() * 4, 10.0 * out.numel() * 5
    if kind == "patch_embed":
        return rec["img"].numel() * 4 + rec["out"].numel() * 2, 2.0 * rec["out"].numel() * 27
    raise KeyError(kind)

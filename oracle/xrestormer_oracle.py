"""CPU oracle for the PromptXRestormer forward.  TEST INFRASTRUCTURE ONLY (same rules as promptir_oracle.py).

Functional fp32 restatement of `net/prompt_xrestormer.py` (kongwanbianjinyu/PromptIR): plain `state_dict` + NCHW image batch ->
what `PromptXRestormer(...).forward` returns.  Pinned by running the real reference next to it (`oracle/make_golden_x.py` ->
`tests/golden/xrestormer_*.npz`); `tests/test_oracle_x.py` re-checks the restatement against those fixtures.
The LayerNorm / MDTA / GDFN / (Un)shuffle pieces are the same arithmetic as net/model.py and are shared with promptir_oracle.py.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from .promptir_oracle import _conv, channel_layernorm, downsample, gdfn, mdta, upsample

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def ocab_core(q: Tensor, k: Tensor, v: Tensor, rel_h: Tensor, rel_w: Tensor, heads: int, dim_head: int = 16, ws: int = 8,
              ows: int = 12) -> Tensor:
    """The attention of OCAB on NCHW q, k, v [B, heads*dim_head, H, W] -> [B, heads*dim_head, H, W].
    prompt_xrestormer.py:215-232 with the relative-position logits of :25-73 written out: queries are the ws x ws windows, keys /
    values the ows x ows window around each (nn.Unfold: ZERO padded at the border -- padded keys still take part in the softmax
    with logit = bias), and
        logit[(x,y),(kr,kc)] = qs . k + qs . rel_width[kc - y + R - 1] + qs . rel_height[kr - x + R - 1],   qs = q * dim_head^-0.5
    with R = ows the relative size (rel_to_abs of :25-37 reduces to that index shift)."""
    b, inner, h, w = q.shape
    pad = (ows - ws) // 2
    nh, nw = h // ws, w // ws
    q = q.view(b, heads, dim_head, nh, ws, nw, ws).permute(0, 3, 5, 1, 4, 6, 2).reshape(b, nh, nw, heads, ws * ws, dim_head)
    q = q * dim_head ** -0.5

    def windows(t):                                        # zero-padded overlapping windows: [b, nh, nw, head, ows*ows, d]
        tp = F.pad(t, (pad, pad, pad, pad))
        u = tp.unfold(2, ows, ws).unfold(3, ows, ws)       # [b, inner, nh, nw, ows, ows]
        return u.reshape(b, heads, dim_head, nh, nw, ows * ows).permute(0, 3, 4, 1, 5, 2)
    kw, vw = windows(k), windows(v)
    logits = q @ kw.transpose(-1, -2)                      # [b, nh, nw, head, 64, 144]
    dev = q.device
    ar_q, ar_k = torch.arange(ws, device=dev), torch.arange(ows, device=dev)
    idx = ar_k.view(1, ows) - ar_q.view(ws, 1) + ows - 1   # [ws, ows]: key coordinate - query coordinate + R - 1
    lw = q @ rel_w.t()                                     # [..., 64, 2R-1]
    lh = q @ rel_h.t()
    qx = torch.arange(ws * ws, device=dev) // ws           # query row
    qy = torch.arange(ws * ws, device=dev) % ws            # query column
    kr = torch.arange(ows * ows, device=dev) // ows
    kc = torch.arange(ows * ows, device=dev) % ows
    bias = torch.gather(lw, -1, idx[qy][:, kc].expand(*lw.shape[:-2], -1, -1)) + \
        torch.gather(lh, -1, idx[qx][:, kr].expand(*lh.shape[:-2], -1, -1))
    attn = torch.softmax(logits + bias, dim=-1)
    out = attn @ vw                                        # [b, nh, nw, head, 64, d]
    return out.view(b, nh, nw, heads, ws, ws, dim_head).permute(0, 3, 6, 1, 4, 2, 5).reshape(b, inner, h, w)


def ocab(x: Tensor, sd: StateDict, p: str, heads: int, dim_head: int = 16, ws: int = 8, overlap: float = 0.5) -> Tensor:
    """Overlapping cross-attention.  prompt_xrestormer.py:209-235."""
    ows = int(ws * overlap) + ws
    q, k, v = _conv(x, sd, p + ".qkv").chunk(3, dim=1)
    out = ocab_core(q, k, v, sd[p + ".rel_pos_emb.rel_height"], sd[p + ".rel_pos_emb.rel_width"], heads, dim_head, ws, ows)
    return _conv(out, sd, p + ".project_out")


def x_block(x: Tensor, sd: StateDict, p: str, channel_heads: int, spatial_heads: int) -> Tensor:
    """prompt_xrestormer.py:255-260."""
    ln = lambda t, n: channel_layernorm(t, sd[f"{p}.{n}.body.weight"], sd.get(f"{p}.{n}.body.bias"))
    x = x + mdta(ln(x, "norm1"), sd, p + ".channel_attn", channel_heads)
    x = x + gdfn(ln(x, "norm2"), sd, p + ".channel_ffn")
    x = x + ocab(ln(x, "norm3"), sd, p + ".spatial_attn", spatial_heads)
    x = x + gdfn(ln(x, "norm4"), sd, p + ".spatial_ffn")
    return x


def x_stage(x: Tensor, sd: StateDict, name: str, depth: int, ch: int, sh: int) -> Tensor:
    for i in range(depth):
        x = x_block(x, sd, f"{name}.{i}", ch, sh)
    return x


def prompt_block(x: Tensor, sd: StateDict, p: str, spatial_heads: int) -> Tensor:
    """prompt_xrestormer.py:343-359.  NOTE align_corners=True here (net/model.py's PromptGenBlock uses False)."""
    h, w = x.shape[-2:]
    emb = x.mean(dim=(-2, -1))
    wts = torch.softmax(F.linear(emb, sd[p + ".linear_layer.weight"], sd[p + ".linear_layer.bias"]), dim=1)
    prompt = torch.einsum("bl,ldst->bdst", wts, sd[p + ".prompt_param"][0])
    prompt = F.interpolate(prompt, (h, w), mode="bilinear", align_corners=True)
    prompt = _conv(prompt, sd, p + ".conv3x3", padding=1)
    x = torch.cat([x, prompt], 1)
    x = x_block(x, sd, p + ".attn", 1, spatial_heads)
    return _conv(x, sd, p + ".conv", padding=1)


def xrestormer_forward(sd: StateDict, img: Tensor, num_blocks: Sequence[int] = (4, 6, 6, 8), num_refinement_blocks: int = 4,
                       channel_heads: Sequence[int] = (1, 2, 4, 8), spatial_heads: Sequence[int] = (2, 2, 3, 4),
                       taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """PromptXRestormer.forward (prompt=True, scale=1).  prompt_xrestormer.py:428-478."""
    def tap(n, t):
        if taps is not None:
            taps[n] = t
        return t
    nb, ch, sh = num_blocks, channel_heads, spatial_heads
    e1 = tap("encoder_level1", x_stage(_conv(img, sd, "patch_embed.proj", padding=1), sd, "encoder_level1", nb[0], ch[0], sh[0]))
    e2 = tap("encoder_level2", x_stage(downsample(e1, sd, "down1_2"), sd, "encoder_level2", nb[1], ch[1], sh[1]))
    e3 = tap("encoder_level3", x_stage(downsample(e2, sd, "down2_3"), sd, "encoder_level3", nb[2], ch[2], sh[2]))
    lat = tap("latent", x_stage(downsample(e3, sd, "down3_4"), sd, "latent", nb[3], ch[3], sh[3]))
    lat = tap("prompt3", prompt_block(lat, sd, "prompt3", 8))
    d3 = _conv(torch.cat([upsample(lat, sd, "up4_3"), e3], 1), sd, "reduce_chan_level3")
    d3 = tap("decoder_level3", x_stage(d3, sd, "decoder_level3", nb[2], ch[2], sh[2]))
    d3 = tap("prompt2", prompt_block(d3, sd, "prompt2", 4))
    d2 = _conv(torch.cat([upsample(d3, sd, "up3_2"), e2], 1), sd, "reduce_chan_level2")
    d2 = tap("decoder_level2", x_stage(d2, sd, "decoder_level2", nb[1], ch[1], sh[1]))
    d2 = tap("prompt1", prompt_block(d2, sd, "prompt1", 2))
    d1 = torch.cat([upsample(d2, sd, "up2_1"), e1], 1)
    d1 = tap("decoder_level1", x_stage(d1, sd, "decoder_level1", nb[0], ch[0], sh[0]))
    d1 = tap("refinement", x_stage(d1, sd, "refinement", num_refinement_blocks, ch[0], sh[0]))
    return _conv(d1, sd, "output", padding=1) + img

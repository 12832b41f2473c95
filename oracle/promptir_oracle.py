"""CPU oracle for the PromptIR restoration forward.  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement of the arithmetic performed by the reference
`net/model.py` (kongwanbianjinyu/PromptIR).  It takes a plain ``state_dict`` (name -> tensor) and an
NCHW image batch and returns what ``PromptIR(decoder=True).forward`` returns.  It exists so the CUDA
path can be checked; nothing in the product package imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import this module.

Parity pin: the reference ships no golden vectors or tests for this path, so the oracle is
pinned by running the *real* reference module next to it in the dev container
(``oracle/make_golden.py``; requires /root/reference) and committing the resulting fixtures under
``tests/golden/``.  ``tests/test_oracle.py`` re-checks the oracle against those fixtures everywhere.

All arithmetic lives in third-party PyTorch (ATen/oneDNN); the reference pins torch==1.8.1
(env.yml:230), this image runs torch 2.11 -- the op semantics used here are identical.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------------
# architecture description (net/model.py:245-320)
# ----------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class StageSpec:
    """One run of TransformerBlocks sharing (dim, heads)."""
    name: str          # state_dict prefix, e.g. "encoder_level1"
    dim: int
    heads: int
    depth: int         # 0 => a single un-indexed block (noise_level*)


@dataclass(frozen=True)
class ArchSpec:
    """Static description of PromptIR(decoder=True) with the reference's hard-coded widths.

    net/model.py:260-320.  The reference hard-codes 64/128/320 prompt widths and 192/224/512 offsets, so
    only dim=48 produces a runnable network."""
    dim: int = 48
    num_blocks: Sequence[int] = (4, 6, 6, 8)
    num_refinement_blocks: int = 4
    heads: Sequence[int] = (1, 2, 4, 8)
    ffn_expansion_factor: float = 2.66
    layernorm_type: str = "WithBias"
    decoder: bool = True

    def hidden(self, c: int) -> int:
        return int(c * self.ffn_expansion_factor)          # net/model.py:86

    def stages(self) -> List[StageSpec]:
        d, nb, hd = self.dim, self.num_blocks, self.heads
        return [
            StageSpec("encoder_level1", d, hd[0], nb[0]),
            StageSpec("encoder_level2", d * 2, hd[1], nb[1]),
            StageSpec("encoder_level3", d * 4, hd[2], nb[2]),
            StageSpec("latent", d * 8, hd[3], nb[3]),
            StageSpec("noise_level3", d * 4 + 512, hd[2], 0),
            StageSpec("decoder_level3", d * 4, hd[2], nb[2]),
            StageSpec("noise_level2", d * 2 + 224, hd[2], 0),
            StageSpec("decoder_level2", d * 2, hd[1], nb[1]),
            StageSpec("noise_level1", d * 2 + 64, hd[2], 0),
            StageSpec("decoder_level1", d * 2, hd[0], nb[0]),
            StageSpec("refinement", d * 2, hd[0], self.num_refinement_blocks),
        ]


# ----------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------
def channel_layernorm(x: Tensor, weight: Tensor, bias: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    """Per-pixel normalisation over the channel axis of an NCHW tensor.

    net/model.py:21-25 (to_3d/to_4d), :60-63 (WithBias), :39-41 (BiasFree).  Variance is the biased
    (population) variance about the mean in both flavours; BiasFree does NOT centre the numerator."""
    mu = x.mean(dim=1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=1, keepdim=True)
    sd = torch.sqrt(var + eps)
    w = weight.view(1, -1, 1, 1)
    if bias is None:
        return x / sd * w
    return (x - mu) / sd * w + bias.view(1, -1, 1, 1)


def _conv(x: Tensor, sd: StateDict, name: str, **kw) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), **kw)


def mdta(x: Tensor, sd: StateDict, p: str, heads: int) -> Tensor:
    """Multi-DConv head transposed attention.  net/model.py:117-138.

    qkv = dw3x3(conv1x1(x)); q,k,v = consecutive thirds of the channel axis; per head (consecutive
    channel groups) q,k are L2-normalised over the H*W axis; logits = (q k^T) * temperature[head];
    softmax over the last (key-channel) axis; out = A v; then project_out."""
    b, c, h, w = x.shape
    qkv = _conv(x, sd, p + ".qkv")
    qkv = _conv(qkv, sd, p + ".qkv_dwconv", padding=1, groups=qkv.shape[1])
    q, k, v = qkv.view(b, 3, heads, c // heads, h * w).unbind(dim=1)      # chunk(3) + head split
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)                 # F.normalize, model.py:127
    k = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)                 # model.py:128
    logits = torch.matmul(q, k.transpose(-1, -2)) * sd[p + ".temperature"].view(1, heads, 1, 1)
    attn = torch.softmax(logits, dim=-1)
    out = torch.matmul(attn, v).reshape(b, c, h, w)
    return _conv(out, sd, p + ".project_out")


def gdfn(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """Gated depthwise-conv feed-forward.  net/model.py:94-99 (exact erf GELU on the first half)."""
    y = _conv(x, sd, p + ".project_in")
    y = _conv(y, sd, p + ".dwconv", padding=1, groups=y.shape[1])
    hdim = y.shape[1] // 2
    gated = F.gelu(y[:, :hdim]) * y[:, hdim:]
    return _conv(gated, sd, p + ".project_out")


def transformer_block(x: Tensor, sd: StateDict, p: str, heads: int) -> Tensor:
    """net/model.py:192-196."""
    n1b = sd.get(p + ".norm1.body.bias")
    n2b = sd.get(p + ".norm2.body.bias")
    x = x + mdta(channel_layernorm(x, sd[p + ".norm1.body.weight"], n1b), sd, p + ".attn", heads)
    x = x + gdfn(channel_layernorm(x, sd[p + ".norm2.body.weight"], n2b), sd, p + ".ffn")
    return x


def run_stage(x: Tensor, sd: StateDict, st: StageSpec) -> Tensor:
    if st.depth == 0:
        return transformer_block(x, sd, st.name, st.heads)
    for i in range(st.depth):
        x = transformer_block(x, sd, f"{st.name}.{i}", st.heads)
    return x


def downsample(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """conv3x3 (n -> n/2) then PixelUnshuffle(2).  net/model.py:164-168."""
    return F.pixel_unshuffle(_conv(x, sd, p + ".body.0", padding=1), 2)


def upsample(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """conv3x3 (n -> 2n) then PixelShuffle(2).  net/model.py:174-178."""
    return F.pixel_shuffle(_conv(x, sd, p + ".body.0", padding=1), 2)


def prompt_weights(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """softmax(Linear(mean_{H,W} x)) -> [B, 5].  net/model.py:228-229."""
    emb = x.mean(dim=(-2, -1))
    return torch.softmax(F.linear(emb, sd[p + ".linear_layer.weight"], sd[p + ".linear_layer.bias"]), dim=1)


def prompt_gen(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """PromptGenBlock.  net/model.py:226-235.  Weighted sum of the learned prompt components, bilinear
    resize (align_corners=False) to the feature size, dense conv3x3."""
    h, w = x.shape[-2:]
    wts = prompt_weights(x, sd, p)                                        # [B, L]
    comps = sd[p + ".prompt_param"][0]                                    # [L, D, S, S]
    prompt = torch.einsum("bl,ldst->bdst", wts, comps)
    prompt = F.interpolate(prompt, (h, w), mode="bilinear")
    return _conv(prompt, sd, p + ".conv3x3", padding=1)


# ----------------------------------------------------------------------------------------------------
# whole network
# ----------------------------------------------------------------------------------------------------
def promptir_forward(sd: StateDict, img: Tensor, arch: ArchSpec = ArchSpec(),
                     taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """PromptIR.forward.  net/model.py:322-380.  `taps`, if given, is filled with named intermediates
    (NCHW fp32) so per-stage parity can be localised."""
    st = {s.name: s for s in arch.stages()}

    def tap(name: str, t: Tensor) -> Tensor:
        if taps is not None:
            taps[name] = t
        return t

    x1 = tap("patch_embed", _conv(img, sd, "patch_embed.proj", padding=1))          # :324
    e1 = tap("encoder_level1", run_stage(x1, sd, st["encoder_level1"]))             # :326
    e2 = tap("encoder_level2", run_stage(downsample(e1, sd, "down1_2"), sd, st["encoder_level2"]))   # :328-330
    e3 = tap("encoder_level3", run_stage(downsample(e2, sd, "down2_3"), sd, st["encoder_level3"]))   # :332-334
    lat = tap("latent", run_stage(downsample(e3, sd, "down3_4"), sd, st["latent"]))                  # :336-337
    if arch.decoder:                                                                                 # :339-343
        lat = torch.cat([lat, tap("prompt3", prompt_gen(lat, sd, "prompt3"))], dim=1)
        lat = run_stage(lat, sd, st["noise_level3"])
        lat = tap("reduce_noise_level3", _conv(lat, sd, "reduce_noise_level3"))
    d3 = torch.cat([upsample(lat, sd, "up4_3"), e3], dim=1)                                          # :346-347
    d3 = _conv(d3, sd, "reduce_chan_level3")                                                         # :348
    d3 = tap("decoder_level3", run_stage(d3, sd, st["decoder_level3"]))                             # :350
    if arch.decoder:                                                                                 # :351-355
        d3 = torch.cat([d3, tap("prompt2", prompt_gen(d3, sd, "prompt2"))], dim=1)
        d3 = run_stage(d3, sd, st["noise_level2"])
        d3 = tap("reduce_noise_level2", _conv(d3, sd, "reduce_noise_level2"))
    d2 = torch.cat([upsample(d3, sd, "up3_2"), e2], dim=1)                                           # :358-359
    d2 = _conv(d2, sd, "reduce_chan_level2")                                                         # :360
    d2 = tap("decoder_level2", run_stage(d2, sd, st["decoder_level2"]))                             # :362
    if arch.decoder:                                                                                 # :363-367
        d2 = torch.cat([d2, tap("prompt1", prompt_gen(d2, sd, "prompt1"))], dim=1)
        d2 = run_stage(d2, sd, st["noise_level1"])
        d2 = tap("reduce_noise_level1", _conv(d2, sd, "reduce_noise_level1"))
    d1 = torch.cat([upsample(d2, sd, "up2_1"), e1], dim=1)                                           # :369-370
    d1 = tap("decoder_level1", run_stage(d1, sd, st["decoder_level1"]))                             # :372
    d1 = tap("refinement", run_stage(d1, sd, st["refinement"]))                                     # :374
    return _conv(d1, sd, "output", padding=1) + img                                                  # :377


# ----------------------------------------------------------------------------------------------------
# callers on either side of the path (demo.py) and metrics
# ----------------------------------------------------------------------------------------------------
def pad_to_multiple(img: Tensor, mult: int = 8):
    """demo.py:17-24.  Reflect-pad bottom/right up to the next multiple (only when not already one)."""
    h, w = img.shape[-2:]
    ph = (h + mult) // mult * mult - h if h % mult else 0
    pw = (w + mult) // mult * mult - w if w % mult else 0
    return F.pad(img, (0, pw, 0, ph), mode="reflect"), h, w


def tile_origins(extent: int, tile: int, overlap: int) -> List[int]:
    """demo.py:32-34: range(0, extent-tile, tile-overlap) + [extent-tile]."""
    return list(range(0, extent - tile, tile - overlap)) + [extent - tile]


def tiled_restore(fn, img: Tensor, tile: int = 128, overlap: int = 32) -> Tensor:
    """demo.py:26-48.  Overlapping tiles, hit-count averaging, clamp to [0,1]."""
    b, c, h, w = img.shape
    tile = min(tile, h, w)
    assert tile % 8 == 0
    acc = torch.zeros_like(img)
    hits = torch.zeros_like(img)
    for y in tile_origins(h, tile, overlap):
        for x in tile_origins(w, tile, overlap):
            acc[..., y:y + tile, x:x + tile] += fn(img[..., y:y + tile, x:x + tile])
            hits[..., y:y + tile, x:x + tile] += 1
    return (acc / hits).clamp(0, 1)


def psnr(a: Tensor, b: Tensor) -> float:
    """PSNR with data_range=1 on clipped tensors (utils/val_utils.py:50-66 uses skimage with data_range=1)."""
    mse = (a.clamp(0, 1).double() - b.clamp(0, 1).double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)


# ----------------------------------------------------------------------------------------------------
# synthetic data (mirrors utils/degradation_utils.py:21-40 for the noise case)
# ----------------------------------------------------------------------------------------------------
def synthetic_batch(b: int, h: int, w: int, seed: int = 1):
    """Returns (degraded, clean) fp32 NCHW in [0,1]; degradation cycles noise15/25/50/rain/haze."""
    g = torch.Generator().manual_seed(seed)
    clean = torch.rand(b, 3, h, w, generator=g)
    clean = F.avg_pool2d(F.pad(clean, (1, 1, 1, 1), mode="reflect"), 3, stride=1)   # mild low-pass
    out = torch.empty_like(clean)
    for i in range(b):
        kind = i % 5
        c = clean[i]
        if kind < 3:
            sigma = (15.0, 25.0, 50.0)[kind]
            n = torch.randn(c.shape, generator=g)
            out[i] = torch.clamp(c * 255.0 + sigma * n, 0, 255).floor() / 255.0       # uint8 quantisation
        elif kind == 3:
            mask = (torch.rand(1, h, w, generator=g) > 0.97).float()
            streak = F.max_pool2d(mask[None], (7, 1), stride=1, padding=(3, 0))[0]
            inten = 0.5 + 0.5 * torch.rand(1, generator=g)
            out[i] = torch.clamp(c + streak * inten, 0, 1)
        else:
            t = 0.3 + 0.6 * torch.rand(1, generator=g)
            a = 0.7 + 0.3 * torch.rand(1, generator=g)
            out[i] = c * t + a * (1 - t)
    return out, clean

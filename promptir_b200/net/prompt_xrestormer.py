"""Drop-in `PromptXRestormer` for kongwanbianjinyu/PromptIR's `net/prompt_xrestormer.py`, executed by the sm_100a engine.

Same constructor arguments, parameter names / shapes / registration order (so seeded construction and `state_dict` round-trips
are identical to the reference, prompt_xrestormer.py:366-425) and the same `forward(inp_img)` (prompt_xrestormer.py:428-478).
The sub-modules only HOLD parameters; `forward` hands the image to `promptir_b200.xengine.XEngine` (no grad) or, with autograd
enabled, to `promptir_b200.xtrain_engine.XTrainEngine` through one autograd node; there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .model import Downsample, FeedForward, LayerNorm, OverlapPatchEmbed, Upsample, _PromptIRFunction


class RelPosEmb(nn.Module):                                 # prompt_xrestormer.py:48-73
    def __init__(self, block_size: int, rel_size: int, dim_head: int):
        super().__init__()
        scale = dim_head ** -0.5
        self.block_size = block_size
        self.rel_height = nn.Parameter(torch.randn(rel_size * 2 - 1, dim_head) * scale)
        self.rel_width = nn.Parameter(torch.randn(rel_size * 2 - 1, dim_head) * scale)


class ChannelAttention(nn.Module):                          # prompt_xrestormer.py:155-165 (MDTA)
    def __init__(self, dim: int, num_heads: int, bias: bool):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=bias)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, 1, 1, groups=dim * 3, bias=bias)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=bias)


class OCAB(nn.Module):                                      # prompt_xrestormer.py:189-208
    def __init__(self, dim: int, window_size: int, overlap_ratio: float, num_heads: int, dim_head: int, bias: bool):
        super().__init__()
        self.num_spatial_heads = num_heads
        self.dim = dim
        self.window_size = window_size
        self.overlap_win_size = int(window_size * overlap_ratio) + window_size
        self.dim_head = dim_head
        self.inner_dim = dim_head * num_heads
        self.scale = dim_head ** -0.5
        self.qkv = nn.Conv2d(dim, self.inner_dim * 3, 1, bias=bias)
        self.project_out = nn.Conv2d(self.inner_dim, dim, 1, bias=bias)
        self.rel_pos_emb = RelPosEmb(window_size, window_size + (self.overlap_win_size - window_size), dim_head)


class TransformerBlock(nn.Module):                          # prompt_xrestormer.py:238-253
    def __init__(self, dim, window_size, overlap_ratio, num_channel_heads, num_spatial_heads, spatial_dim_head, ffn_expansion_factor, bias,
                 LayerNorm_type):
        super().__init__()
        self.spatial_attn = OCAB(dim, window_size, overlap_ratio, num_spatial_heads, spatial_dim_head, bias)
        self.channel_attn = ChannelAttention(dim, num_channel_heads, bias)
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        self.norm3 = LayerNorm(dim, LayerNorm_type)
        self.norm4 = LayerNorm(dim, LayerNorm_type)
        self.channel_ffn = FeedForward(dim, ffn_expansion_factor, bias)
        self.spatial_ffn = FeedForward(dim, ffn_expansion_factor, bias)


class PromptBlock(nn.Module):                               # prompt_xrestormer.py:322-341
    def __init__(self, window_size, overlap_ratio, num_channel_heads, num_spatial_heads, spatial_dim_head, ffn_expansion_factor, bias,
                 LayerNorm_type, prompt_dim=128, prompt_len=5, prompt_size=96, lin_dim=192):
        super().__init__()
        self.prompt_param = nn.Parameter(torch.rand(1, prompt_len, prompt_dim, prompt_size, prompt_size))
        self.linear_layer = nn.Linear(lin_dim, prompt_len)
        self.conv3x3 = nn.Conv2d(prompt_dim, prompt_dim, 3, 1, 1, bias=False)
        self.attn = TransformerBlock(lin_dim + prompt_dim, window_size, overlap_ratio, num_channel_heads, num_spatial_heads, spatial_dim_head,
                                     ffn_expansion_factor, bias, LayerNorm_type)
        self.conv = nn.Conv2d(prompt_dim + lin_dim, lin_dim, 3, 1, 1, bias=False)


class PromptXRestormer(nn.Module):
    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4, channel_heads=[1, 2, 4, 8],
                 spatial_heads=[2, 2, 3, 4], overlap_ratio=[0.5, 0.5, 0.5, 0.5], window_size=8, spatial_dim_head=16, bias=False,
                 ffn_expansion_factor=2.66, LayerNorm_type="WithBias", dual_pixel_task=False, scale=1, prompt=True):
        super().__init__()
        self.scale = scale

        def stage(n, d, lvl):
            return nn.Sequential(*[TransformerBlock(d, window_size, overlap_ratio[lvl], channel_heads[lvl], spatial_heads[lvl], spatial_dim_head,
                                                    ffn_expansion_factor, bias, LayerNorm_type) for _ in range(n)])
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.encoder_level1 = stage(num_blocks[0], dim, 0)
        self.down1_2 = Downsample(dim)
        self.encoder_level2 = stage(num_blocks[1], dim * 2, 1)
        self.down2_3 = Downsample(dim * 2)
        self.encoder_level3 = stage(num_blocks[2], dim * 4, 2)
        self.down3_4 = Downsample(dim * 4)
        self.latent = stage(num_blocks[3], dim * 8, 3)
        self.up4_3 = Upsample(dim * 8)
        self.reduce_chan_level3 = nn.Conv2d(dim * 8, dim * 4, 1, bias=bias)
        self.decoder_level3 = stage(num_blocks[2], dim * 4, 2)
        self.up3_2 = Upsample(dim * 4)
        self.reduce_chan_level2 = nn.Conv2d(dim * 4, dim * 2, 1, bias=bias)
        self.decoder_level2 = stage(num_blocks[1], dim * 2, 1)
        self.up2_1 = Upsample(dim * 2)
        self.decoder_level1 = stage(num_blocks[0], dim * 2, 0)
        self.refinement = stage(num_refinement_blocks, dim * 2, 0)
        self.output = nn.Conv2d(dim * 2, out_channels, 3, 1, 1, bias=bias)
        self.prompt = prompt
        if prompt:                                          # hard-coded widths, prompt_xrestormer.py:415-426 (dim = 48 only)
            kw = dict(window_size=8, overlap_ratio=0.5, num_channel_heads=1, spatial_dim_head=spatial_dim_head,
                      ffn_expansion_factor=ffn_expansion_factor, bias=bias, LayerNorm_type=LayerNorm_type)
            self.prompt1 = PromptBlock(prompt_dim=64, prompt_len=5, prompt_size=64, lin_dim=96, num_spatial_heads=2, **kw)
            self.prompt2 = PromptBlock(prompt_dim=128, prompt_len=5, prompt_size=32, lin_dim=192, num_spatial_heads=4, **kw)
            self.prompt3 = PromptBlock(prompt_dim=320, prompt_len=5, prompt_size=16, lin_dim=384, num_spatial_heads=8, **kw)

        self.layernorm_type = LayerNorm_type
        self.window_size = window_size
        self.compute_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("PROMPTIR_B200_DTYPE", "bf16")]
        self.use_cuda_graph = os.environ.get("PROMPTIR_B200_GRAPH", "1") != "0"
        self._engines: Dict[Tuple, object] = {}
        self._train_engine = None
        self.grad_scale = None                              # static loss scale of the 16-bit backward (None: 1 for bf16, 65536 for fp16)

    def train_engine_for(self, batch: int, height: int, width: int, device: torch.device, input_grad: bool = False):
        from ..xtrain_engine import XTrainEngine
        key = (batch, height, width, str(device), self.compute_dtype, input_grad, self.grad_scale)
        if self._train_engine is None or self._train_engine[0] != key:
            self._train_engine = None
            self._train_engine = (key, XTrainEngine(self, batch, height, width, device, self.compute_dtype, grad_scale=self.grad_scale,
                                                    input_grad=input_grad))
        return self._train_engine[1]

    def engine_for(self, batch: int, height: int, width: int, device: torch.device):
        from ..xengine import XEngine
        key = (batch, height, width, str(device), self.compute_dtype)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 2:
                self._engines.pop(next(iter(self._engines)))
            eng = XEngine(self, batch, height, width, device, self.compute_dtype)
            self._engines[key] = eng
        return eng

    def _apply(self, fn, *a, **k):
        self._engines = {}
        self._train_engine = None
        return super()._apply(fn, *a, **k)

    def forward(self, inp_img: torch.Tensor) -> torch.Tensor:              # prompt_xrestormer.py:428
        if self.scale != 1:
            raise NotImplementedError("promptir_b200.PromptXRestormer: scale > 1 (the super-resolution pre-resize) is not built")
        if not self.prompt:
            raise NotImplementedError("promptir_b200.PromptXRestormer: prompt=False is not built (every config of the reference uses prompts)")
        if inp_img.dim() != 4:
            raise ValueError(f"expected a [B, C, H, W] image batch, got {tuple(inp_img.shape)}")
        if not inp_img.is_cuda:
            raise RuntimeError("promptir_b200.PromptXRestormer runs on a B200 (sm_100a) only; got a CPU tensor and there is no CPU fallback")
        b, c, h, w = inp_img.shape
        m = 8 * self.window_size
        if h % m or w % m:
            raise RuntimeError(f"height and width must be multiples of {m} (three 2x downsamples and {self.window_size}x{self.window_size} "
                               f"attention windows at every level), got {h}x{w}; the reference fails in rearrange for such sizes too")
        x = inp_img if (inp_img.dtype == torch.float32 and inp_img.is_contiguous()) else inp_img.float().contiguous()
        if torch.is_grad_enabled() and (inp_img.requires_grad or any(p.requires_grad for p in self.parameters())):
            eng = self.train_engine_for(b, h, w, inp_img.device, input_grad=inp_img.requires_grad)
            out = _PromptIRFunction.apply(self, eng, x, *[p for _, p in self.named_parameters()])
            return out if inp_img.dtype == torch.float32 else out.to(inp_img.dtype)
        eng = self.engine_for(b, h, w, inp_img.device)
        out = eng.run(x, use_graph=self.use_cuda_graph)
        return out if inp_img.dtype == torch.float32 else out.to(inp_img.dtype)

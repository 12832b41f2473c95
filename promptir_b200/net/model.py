"""Drop-in `PromptIR` for kongwanbianjinyu/PromptIR's `net/model.py`, executed by the sm_100a engine.

Same constructor arguments, parameter names/shapes/registration order (so seeded construction and
`state_dict` round-trips are identical to the reference, net/model.py:245-320) and the same
`forward(inp_img, noise_emb=None)` (net/model.py:322).  The sub-modules below only HOLD parameters in the
reference's hierarchy; none of them computes anything.  `PromptIR.forward` hands the image to
`promptir_b200.engine.Engine`, which runs the hand-written CUDA kernels through the C ABI
(include/promptir_b200.h).  There is no PyTorch/CPU fallback: a CPU tensor or a missing library raises.

Compute type: 16-bit storage/operands with fp32 accumulation.  `compute_dtype` (attribute, or the
PROMPTIR_B200_DTYPE=bf16|fp16 environment variable; default bf16) selects bfloat16 or float16.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------------------------------
# parameter containers (reference hierarchy; registration order matters for RNG-identical init)
# ----------------------------------------------------------------------------------------------------
class BiasFree_LayerNorm(nn.Module):                       # model.py:27-41
    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class WithBias_LayerNorm(nn.Module):                       # model.py:47-63
    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class LayerNorm(nn.Module):                                # model.py:66-76
    def __init__(self, dim: int, LayerNorm_type: str):
        super().__init__()
        self.body = BiasFree_LayerNorm(dim) if LayerNorm_type == "BiasFree" else WithBias_LayerNorm(dim)


class FeedForward(nn.Module):                              # model.py:82-92
    def __init__(self, dim: int, ffn_expansion_factor: float, bias: bool):
        super().__init__()
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = nn.Conv2d(dim, hidden * 2, 1, bias=bias)
        self.dwconv = nn.Conv2d(hidden * 2, hidden * 2, 3, 1, 1, groups=hidden * 2, bias=bias)
        self.project_out = nn.Conv2d(hidden, dim, 1, bias=bias)


class Attention(nn.Module):                                # model.py:105-113
    def __init__(self, dim: int, num_heads: int, bias: bool):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=bias)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, 1, 1, groups=dim * 3, bias=bias)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=bias)


class TransformerBlock(nn.Module):                         # model.py:183-190
    def __init__(self, dim: int, num_heads: int, ffn_expansion_factor: float, bias: bool, LayerNorm_type: str):
        super().__init__()
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        self.attn = Attention(dim, num_heads, bias)
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)


class OverlapPatchEmbed(nn.Module):                        # model.py:202-206
    def __init__(self, in_c: int = 3, embed_dim: int = 48, bias: bool = False):
        super().__init__()
        self.proj = nn.Conv2d(in_c, embed_dim, 3, 1, 1, bias=bias)


class resblock(nn.Module):                                 # model.py:142-155 (never instantiated by the reference; kept for import parity)
    def __init__(self, dim: int):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, bias=False), nn.PReLU(),
                                  nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, bias=False))

    def forward(self, x):
        # Not on the hot path (no caller in the reference): conv-PReLU-conv residual through torch, like model.py:153-155
        return self.body(x) + x


class Downsample(nn.Module):                               # model.py:160-165
    def __init__(self, n_feat: int):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat // 2, 3, 1, 1, bias=False), nn.PixelUnshuffle(2))


class Upsample(nn.Module):                                 # model.py:170-175
    def __init__(self, n_feat: int):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat * 2, 3, 1, 1, bias=False), nn.PixelShuffle(2))


class PromptGenBlock(nn.Module):                           # model.py:218-223
    def __init__(self, prompt_dim: int = 128, prompt_len: int = 5, prompt_size: int = 96, lin_dim: int = 192):
        super().__init__()
        self.prompt_param = nn.Parameter(torch.rand(1, prompt_len, prompt_dim, prompt_size, prompt_size))
        self.linear_layer = nn.Linear(lin_dim, prompt_len)
        self.conv3x3 = nn.Conv2d(prompt_dim, prompt_dim, 3, 1, 1, bias=False)


def _blocks(n: int, dim: int, heads: int, f: float, bias: bool, ln: str) -> nn.Sequential:
    return nn.Sequential(*[TransformerBlock(dim, heads, f, bias, ln) for _ in range(n)])


# ----------------------------------------------------------------------------------------------------
class _PromptIRFunction(torch.autograd.Function):
    """Autograd node of the whole network: forward = the training launch program, backward = the hand-written backward program
    (promptir_b200/train_engine.py).  Replaces the ~5000 ATen backward kernels autograd runs for net/model.py under train.py:41-46."""

    @staticmethod
    def forward(ctx, module, eng, img, *params):
        out = eng.forward(img, use_graph=module.use_cuda_graph)
        eng.generation += 1
        ctx.module, ctx.eng, ctx.generation = module, eng, eng.generation
        ctx.need_img = img.requires_grad
        ctx.need = [p.requires_grad for p in params]
        return out

    @staticmethod
    def backward(ctx, d_out):
        eng, module = ctx.eng, ctx.module
        if eng.generation != ctx.generation:
            raise RuntimeError("promptir_b200: the activations kept for this backward were overwritten by a later differentiable "
                               "forward of the same shape; call backward() before the next forward (as train.py does)")
        eng.backward(d_out.contiguous().float(), use_graph=module.use_cuda_graph)
        grads = []
        for (name, p), need in zip(module.named_parameters(), ctx.need):
            # parameters the forward never reads (chnl_reduce*, reduce_noise_channel_*: model.py:271-287) get no gradient, like autograd
            g = eng.grads[name] if (need and name in eng.live_params) else None
            # p.grad may already BE this view (ddp.attach_flat_grads): AccumulateGrad would then compute p.grad += p.grad on one
            # buffer; any other existing .grad would keep a reference into a buffer the next backward overwrites.  A clone is safe.
            if g is not None and p.grad is not None:
                g = g.clone()
            grads.append(g)
        return (None, None, eng.d_img.clone() if ctx.need_img else None, *grads)


class PromptIR(nn.Module):
    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                 heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias", decoder=False):
        super().__init__()
        f, ln = ffn_expansion_factor, LayerNorm_type
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.decoder = decoder
        if self.decoder:
            self.prompt1 = PromptGenBlock(prompt_dim=64, prompt_len=5, prompt_size=64, lin_dim=96)
            self.prompt2 = PromptGenBlock(prompt_dim=128, prompt_len=5, prompt_size=32, lin_dim=192)
            self.prompt3 = PromptGenBlock(prompt_dim=320, prompt_len=5, prompt_size=16, lin_dim=384)
        # constructed by the reference but never used by its forward (model.py:271-287); kept for state_dict parity
        self.chnl_reduce1 = nn.Conv2d(64, 64, 1, bias=bias)
        self.chnl_reduce2 = nn.Conv2d(128, 128, 1, bias=bias)
        self.chnl_reduce3 = nn.Conv2d(320, 256, 1, bias=bias)
        self.reduce_noise_channel_1 = nn.Conv2d(dim + 64, dim, 1, bias=bias)
        self.encoder_level1 = _blocks(num_blocks[0], dim, heads[0], f, bias, ln)
        self.down1_2 = Downsample(dim)
        self.reduce_noise_channel_2 = nn.Conv2d(int(dim * 2) + 128, int(dim * 2), 1, bias=bias)
        self.encoder_level2 = _blocks(num_blocks[1], int(dim * 2), heads[1], f, bias, ln)
        self.down2_3 = Downsample(int(dim * 2))
        self.reduce_noise_channel_3 = nn.Conv2d(int(dim * 4) + 256, int(dim * 4), 1, bias=bias)
        self.encoder_level3 = _blocks(num_blocks[2], int(dim * 4), heads[2], f, bias, ln)
        self.down3_4 = Downsample(int(dim * 4))
        self.latent = _blocks(num_blocks[3], int(dim * 8), heads[3], f, bias, ln)
        self.up4_3 = Upsample(int(dim * 4))
        self.reduce_chan_level3 = nn.Conv2d(int(dim * 2) + 192, int(dim * 4), 1, bias=bias)
        self.noise_level3 = TransformerBlock(int(dim * 4) + 512, heads[2], f, bias, ln)
        self.reduce_noise_level3 = nn.Conv2d(int(dim * 4) + 512, int(dim * 4), 1, bias=bias)
        self.decoder_level3 = _blocks(num_blocks[2], int(dim * 4), heads[2], f, bias, ln)
        self.up3_2 = Upsample(int(dim * 4))
        self.reduce_chan_level2 = nn.Conv2d(int(dim * 4), int(dim * 2), 1, bias=bias)
        self.noise_level2 = TransformerBlock(int(dim * 2) + 224, heads[2], f, bias, ln)
        self.reduce_noise_level2 = nn.Conv2d(int(dim * 2) + 224, int(dim * 4), 1, bias=bias)
        self.decoder_level2 = _blocks(num_blocks[1], int(dim * 2), heads[1], f, bias, ln)
        self.up2_1 = Upsample(int(dim * 2))
        self.noise_level1 = TransformerBlock(int(dim * 2) + 64, heads[2], f, bias, ln)
        self.reduce_noise_level1 = nn.Conv2d(int(dim * 2) + 64, int(dim * 2), 1, bias=bias)
        self.decoder_level1 = _blocks(num_blocks[0], int(dim * 2), heads[0], f, bias, ln)
        self.refinement = _blocks(num_refinement_blocks, int(dim * 2), heads[0], f, bias, ln)
        self.output = nn.Conv2d(int(dim * 2), out_channels, 3, 1, 1, bias=bias)

        self.layernorm_type = ln
        self.compute_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("PROMPTIR_B200_DTYPE", "bf16")]
        self.use_cuda_graph = os.environ.get("PROMPTIR_B200_GRAPH", "1") != "0"
        self._engines: Dict[Tuple, object] = {}
        self._train_engine = None
        self.grad_scale: Optional[float] = None             # static loss scale of the 16-bit backward (None: 1 for bf16, 65536 for fp16)

    # -- engine cache ------------------------------------------------------------------------------
    def engine_for(self, batch: int, height: int, width: int, device: torch.device):
        from ..engine import Engine, SplitEngine
        key = (batch, height, width, str(device), self.compute_dtype)
        eng = self._engines.get(key)
        if eng is not None and eng.params_moved():          # p.data = ..., load_state_dict(assign=True), EMA swap: raw pointers are stale
            del self._engines[key]
            eng = None
        if eng is None:
            if len(self._engines) >= 4:                     # bound the workspace held by stale shapes
                self._engines.pop(next(iter(self._engines)))
            # PROMPTIR_B200_SPLIT=1: run batches of >= 4 images as two half-batch branches of one CUDA graph (measured gain on
            # B200 at B=16, 256x256: 0.7 %, so it is off by default)
            split = self.use_cuda_graph and batch >= 4 and os.environ.get("PROMPTIR_B200_SPLIT", "0") == "1"
            eng = (SplitEngine if split else Engine)(self, batch, height, width, device, self.compute_dtype)
            self._engines[key] = eng
        return eng

    def train_engine_for(self, batch: int, height: int, width: int, device: torch.device, input_grad: bool = False):
        """The forward+backward program for this shape (one is kept: it holds every block's activations)."""
        from ..train_engine import TrainEngine
        key = (batch, height, width, str(device), self.compute_dtype, input_grad, self.grad_scale)
        if self._train_engine is None or self._train_engine[0] != key or self._train_engine[1].params_moved():
            self._train_engine = None                       # free the old arena before building the new one
            self._train_engine = (key, TrainEngine(self, batch, height, width, device, self.compute_dtype,
                                                   grad_scale=self.grad_scale, input_grad=input_grad))
        return self._train_engine[1]

    def _apply(self, fn, *a, **k):                          # .to()/.cuda() moves parameters -> caches are stale
        self._engines = {}
        self._train_engine = None
        return super()._apply(fn, *a, **k)

    def forward(self, inp_img: torch.Tensor, noise_emb=None) -> torch.Tensor:   # model.py:322 (noise_emb is ignored there too)
        if not self.decoder:
            # The reference crashes here as well: model.py:346 feeds the 384-channel latent to Upsample(192).
            raise RuntimeError("PromptIR(decoder=False) is not runnable (reference model.py:346 channel mismatch); "
                               "use decoder=True as every caller of the reference does")
        if inp_img.dim() != 4:
            raise ValueError(f"expected a [B, C, H, W] image batch, got {tuple(inp_img.shape)}")
        if not inp_img.is_cuda:
            raise RuntimeError("promptir_b200.PromptIR runs on a B200 (sm_100a) only; got a CPU tensor and there is "
                               "no CPU fallback. Move the module and the input to cuda.")
        b, c, h, w = inp_img.shape
        if h % 8 or w % 8:
            raise RuntimeError(f"pixel_unshuffle expects height and width to be divisible by 8 across the three "
                               f"downsamples, got {h}x{w} (same constraint as the reference; see demo.py pad_input)")
        x = inp_img if (inp_img.dtype == torch.float32 and inp_img.is_contiguous()) else inp_img.float().contiguous()
        if torch.is_grad_enabled() and (inp_img.requires_grad or any(p.requires_grad for p in self.parameters())):
            # differentiable call (train.py:41): the training program keeps activations and autograd gets a hand-written backward
            eng = self.train_engine_for(b, h, w, inp_img.device, input_grad=inp_img.requires_grad)
            params = [p for _, p in self.named_parameters()]
            out = _PromptIRFunction.apply(self, eng, x, *params)
            return out if inp_img.dtype == torch.float32 else out.to(inp_img.dtype)
        eng = self.engine_for(b, h, w, inp_img.device)
        out = eng.run(x, use_graph=self.use_cuda_graph)
        return out if inp_img.dtype == torch.float32 else out.to(inp_img.dtype)

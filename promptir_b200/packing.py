"""Weight repacking for the sm_100a kernels (host-side plumbing; pure torch, device agnostic).

The fp32 nn.Parameters stay the canonical storage (state_dict / optimizers see them unchanged); what is built
here are derived, read-only caches in the layouts the kernels consume:

  * pointwise (1x1) weights  -> 16-bit K-major [N, Kpad], Kpad = ceil(K/64)*64 zero padded; when a LayerNorm
    feeds the conv its gamma is multiplied into the columns and two fp32 vectors carry the rest:
        W.LN(x) = rstd * ( (W*gamma).x  -  mu * s )  +  t,   s = rowsum(round16(W*gamma)),  t = W.beta (+ bias)
    `s` is taken from the ROUNDED weights so the mean term cancels exactly.
  * dense 3x3 weights        -> 16-bit [N, 9*Kpad], tap-major (tap = ky*3 + kx), each tap zero padded to Kpad
  * depthwise 3x3 weights    -> 16-bit [9, Cin] tap-major
  * GDFN hidden width h = int(2.66*C) is padded to hp = ceil(h/8)*8; padded channels have zero weights
    everywhere so they carry exact zeros through dwconv, gate and project_out.
  * prompt components        -> fp32 [L, S, S, D] (channels last)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

Tensor = torch.Tensor


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def kpad_of(k: int) -> int:
    return round_up(k, 64)


def pack_pointwise(w: Tensor, dtype: torch.dtype, *, gamma: Optional[Tensor] = None, beta: Optional[Tensor] = None,
                   bias: Optional[Tensor] = None, k_total: Optional[int] = None,
                   row_map: Optional[Tensor] = None, n_total: Optional[int] = None,
                   col_map: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """w: [N, K] (or [N, K, 1, 1]) fp32 -> (w16 [Ntot, Kpad], ln_s [Ntot] | None, vec_t [Ntot] | None).

    row_map / n_total scatter the N output rows into a larger padded row space (GDFN project_in);
    col_map / k_total scatter the K input columns into a larger padded column space (GDFN project_out)."""
    w = w.detach().reshape(w.shape[0], -1).float()
    n, k = w.shape
    wg = w * gamma.detach().float().view(1, -1) if gamma is not None else w
    t = None
    if beta is not None:
        t = w @ beta.detach().float()
    if bias is not None:
        t = bias.detach().float() if t is None else t + bias.detach().float()
    kt = k if k_total is None else k_total
    nt = n if n_total is None else n_total
    rows = torch.arange(n, device=w.device) if row_map is None else row_map
    cols = torch.arange(k, device=w.device) if col_map is None else col_map
    full = torch.zeros(nt, kpad_of(kt), dtype=torch.float32, device=w.device)
    full[rows[:, None], cols[None, :]] = wg
    w16 = full.to(dtype).contiguous()
    ln_s = w16.float().sum(dim=1).contiguous() if gamma is not None else None
    vec_t = None
    if t is not None:
        vec_t = torch.zeros(nt, dtype=torch.float32, device=w.device)
        vec_t[rows] = t
    return w16, ln_s, vec_t


def pack_conv3x3(w: Tensor, dtype: torch.dtype) -> Tensor:
    """w: [N, Cin, 3, 3] fp32 -> [N, 9*Kpad] 16-bit, column = tap*Kpad + cin, tap = ky*3 + kx."""
    w = w.detach().float()
    n, cin = w.shape[:2]
    kp = kpad_of(cin)
    full = torch.zeros(n, 9, kp, dtype=torch.float32, device=w.device)
    full[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(n, 9, cin)
    return full.reshape(n, 9 * kp).to(dtype).contiguous()


def pack_depthwise(w: Tensor, dtype: torch.dtype, *, chan_map: Optional[Tensor] = None, c_total: Optional[int] = None) -> Tensor:
    """w: [C, 1, 3, 3] fp32 -> [9, Ctot] 16-bit tap-major (optionally scattered into a padded channel space)."""
    w = w.detach().float().reshape(w.shape[0], 9)
    c = w.shape[0]
    ct = c if c_total is None else c_total
    idx = torch.arange(c, device=w.device) if chan_map is None else chan_map
    full = torch.zeros(9, ct, dtype=torch.float32, device=w.device)
    full[:, idx] = w.t()
    return full.to(dtype).contiguous()


def scatter_vec(v: Optional[Tensor], chan_map: Tensor, c_total: int) -> Optional[Tensor]:
    if v is None:
        return None
    out = torch.zeros(c_total, dtype=torch.float32, device=v.device)
    out[chan_map] = v.detach().float()
    return out


def gdfn_maps(h: int, device) -> Tuple[int, Tensor]:
    """hidden width h -> (hp, map of the 2h project_in/dwconv channels into the [x1 | x2] padded space of 2*hp)."""
    hp = round_up(h, 8)
    ar = torch.arange(h, device=device)
    return hp, torch.cat([ar, ar + hp])


def pack_prompt(p: Tensor) -> Tensor:
    """prompt_param [1, L, D, S, S] -> fp32 [L, S, S, D]."""
    return p.detach().float()[0].permute(0, 2, 3, 1).contiguous()

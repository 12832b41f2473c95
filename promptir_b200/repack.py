"""Derived weight caches of an engine, rebuilt on the device by ONE pir_repack launch.

`packing.py` states the cache layouts in torch (what the kernels expect, and what the CPU wiring tests run).  A `Packer` records
the same requests as `PirPackJob` rows (include/promptir_b200.h) instead of computing them: destinations are carved out of a few
zero-initialised arenas, the job table lives in device memory, and `run()` is a single kernel launch that reads the live fp32
parameters -- at engine construction, after `optimizer.step()` (train.py:52-56) and after `load_state_dict`.  On a CPU device
(tests) every request is evaluated immediately with `packing.py` and `run()` replays those torch packers.

Requests (all return tensors that stay valid for the engine's lifetime):
    pointwise(w, gamma=, beta=, bias=, transpose=, rows=(h, hp), cols=(h, hp), n_total=, k_total=) -> (w16, ln_s | None, vec_t | None)
    conv3x3(w, transpose_flip=)      -> w16 [N, 9 * Kpad]
    depthwise(w, dtype, flip=, split=(h, hp), c_total=) -> [9, Ctot]
    vec(v, split=(h, hp), total=)    -> fp32 [total] | None
    prompt(p)                        -> fp32 [L, S, S, D]
    f32(p, shape=)                   -> the parameter itself (fp32, contiguous: no copy), optionally reshaped
`rows` / `cols` / `split` = (h, hp): the GDFN [x1 | x2] padded channel space of packing.gdfn_maps (index h + j -> hp + j).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import torch

from . import packing

Tensor = torch.Tensor


class Packer:
    CHUNK = 32 << 20          # bytes per arena chunk

    def __init__(self, device, verify: bool = False):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.verify = verify and self.cuda
        self.checks: List[Tuple[str, Tensor, Tensor]] = []       # (what, device result, torch reference) when verify
        self._jobs: List[dict] = []
        self._torch: List[Callable[[], None]] = []               # CPU: torch refreshers
        self._chunks: List[Tensor] = []
        self._used = 0
        self._table = None                                         # (jobs_dev, first_row_dev, n_jobs, n_rows, ptr fingerprint)
        self._params: List[Tensor] = []                            # every source tensor (pointer fingerprint)

    # ------------------------------------------------------------------------------------------------
    def _alloc(self, shape, dtype: torch.dtype) -> Tensor:
        n = 1
        for v in shape:
            n *= int(v)
        nbytes = (n * torch.empty((), dtype=dtype).element_size() + 255) // 256 * 256
        if not self._chunks or self._used + nbytes > self._chunks[-1].numel():
            self._chunks.append(torch.zeros(max(self.CHUNK, nbytes), dtype=torch.uint8, device=self.device))
            self._used = 0
        buf = self._chunks[-1][self._used:self._used + n * torch.empty((), dtype=dtype).element_size()]
        self._used += nbytes
        return buf.view(dtype).view(*shape)

    @staticmethod
    def _src(t: Tensor) -> Tensor:
        """The parameter object itself is kept (not a detached alias), so a later `p.data = ...` shows up in data_ptr()."""
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("promptir_b200: parameters must be contiguous fp32 tensors (they are the canonical storage)")
        return t

    def _job(self, **kw) -> None:
        self._jobs.append(kw)
        for key in ("src", "gamma", "beta", "bias"):
            if kw.get(key) is not None:
                self._params.append(kw[key])
        self._table = None

    def _check(self, what: str, got, ref) -> None:
        if self.verify:
            for g, r in zip(got, ref):
                if g is not None:
                    self.checks.append((what, g, r))

    # ------------------------------------------------------------------------------------------------
    def pointwise(self, w: Tensor, dtype: torch.dtype, *, gamma: Optional[Tensor] = None, beta: Optional[Tensor] = None,
                  bias: Optional[Tensor] = None, transpose: bool = False, rows: Optional[Tuple[int, int]] = None,
                  cols: Optional[Tuple[int, int]] = None, n_total: Optional[int] = None, k_total: Optional[int] = None):
        """w: stored [N, K(,1,1)].  transpose: the packed matrix is w^T (logical [K, N]; gamma then scales the logical ROWS)."""
        n_st, k_st = w.shape[0], w[0].numel()
        n, k = (k_st, n_st) if transpose else (n_st, k_st)
        assert not (transpose and (beta is not None or bias is not None))

        def torch_ref():
            w2 = w.detach().reshape(n_st, k_st).float()
            dev = w2.device
            if transpose:
                if gamma is not None:
                    w2 = w2 * gamma.detach().float().view(1, -1)
                w2, g = w2.t(), None
            else:
                g = gamma
            rmap = None if rows is None else torch.cat([torch.arange(rows[0], device=dev), torch.arange(n - rows[0], device=dev) + rows[1]])
            cmap = None if cols is None else torch.cat([torch.arange(cols[0], device=dev), torch.arange(k - cols[0], device=dev) + cols[1]])
            if rows is None and n_total is not None:
                rmap = torch.arange(n, device=dev)
            out = packing.pack_pointwise(w2, dtype, gamma=g, beta=beta, bias=bias, k_total=k_total, row_map=rmap, n_total=n_total, col_map=cmap)
            if transpose:
                return (out[0], None, None)
            return out

        if not self.cuda:
            cur = list(torch_ref())
            self._torch.append(lambda: [d.copy_(s) for d, s in zip(cur, torch_ref()) if d is not None])
            return tuple(cur)
        nt = n if n_total is None else n_total
        kp = packing.kpad_of(k if k_total is None else k_total)
        w16 = self._alloc((nt, kp), dtype)
        want_s = gamma is not None and not transpose
        want_t = (beta is not None or bias is not None) and not transpose
        ln_s = self._alloc((nt,), torch.float32) if want_s else None
        vec_t = self._alloc((nt,), torch.float32) if want_t else None
        self._job(kind=0, dst_dtype=dtype, n=n, k=k, transpose=int(transpose), flip=0, n_total=nt, k_pad=kp,
                  row_split=rows[0] if rows else 0, row_hp=rows[1] if rows else 0, col_split=cols[0] if cols else 0,
                  col_hp=cols[1] if cols else 0, gamma_axis=0 if gamma is None else (2 if transpose else 1),
                  src=self._src(w), gamma=None if gamma is None else self._src(gamma), beta=None if beta is None else self._src(beta),
                  bias=None if bias is None else self._src(bias), dst=w16, ln_s=ln_s, vec_t=vec_t)
        if self.verify:
            self._check("pointwise", (w16, ln_s, vec_t), torch_ref())
        return w16, ln_s, vec_t

    def conv3x3(self, w: Tensor, dtype: torch.dtype, transpose_flip: bool = False) -> Tensor:
        """w: stored [N, Cin, 3, 3].  transpose_flip: the transposed convolution's weights, w.transpose(0, 1).flip(2, 3)."""
        def torch_ref():
            src = w.detach().transpose(0, 1).flip(2, 3) if transpose_flip else w
            return packing.pack_conv3x3(src, dtype)

        if not self.cuda:
            cur = torch_ref()
            self._torch.append(lambda: cur.copy_(torch_ref()))
            return cur
        n, cin = (w.shape[1], w.shape[0]) if transpose_flip else (w.shape[0], w.shape[1])
        kp = packing.kpad_of(cin)
        w16 = self._alloc((n, 9 * kp), dtype)
        self._job(kind=1, dst_dtype=dtype, n=n, k=cin, transpose=int(transpose_flip), flip=int(transpose_flip), n_total=n, k_pad=kp,
                  src=self._src(w), dst=w16)
        if self.verify:
            self._check("conv3x3", (w16,), (torch_ref(),))
        return w16

    def depthwise(self, w: Tensor, dtype: torch.dtype, *, flip: bool = False, split: Optional[Tuple[int, int]] = None,
                  c_total: Optional[int] = None) -> Tensor:
        def torch_ref():
            src = w.detach().flip(2, 3) if flip else w
            cmap = None
            if split is not None:
                dev = w.device
                cmap = torch.cat([torch.arange(split[0], device=dev), torch.arange(w.shape[0] - split[0], device=dev) + split[1]])
            return packing.pack_depthwise(src, dtype, chan_map=cmap, c_total=c_total)

        if not self.cuda:
            cur = torch_ref()
            self._torch.append(lambda: cur.copy_(torch_ref()))
            return cur
        c = w.shape[0]
        ct = c if c_total is None else c_total
        out = self._alloc((9, ct), dtype)
        self._job(kind=2, dst_dtype=dtype, n=c, k=9, flip=int(flip), n_total=ct, k_pad=0, row_split=split[0] if split else 0,
                  row_hp=split[1] if split else 0, src=self._src(w), dst=out)
        if self.verify:
            self._check("depthwise", (out,), (torch_ref(),))
        return out

    def vec(self, v: Optional[Tensor], split: Tuple[int, int], total: int) -> Optional[Tensor]:
        if v is None:
            return None

        def torch_ref():
            dev = v.device
            cmap = torch.cat([torch.arange(split[0], device=dev), torch.arange(v.numel() - split[0], device=dev) + split[1]])
            return packing.scatter_vec(v, cmap, total)

        if not self.cuda:
            cur = torch_ref()
            self._torch.append(lambda: cur.copy_(torch_ref()))
            return cur
        out = self._alloc((total,), torch.float32)
        self._job(kind=3, dst_dtype=torch.float32, n=v.numel(), k=1, n_total=total, k_pad=0, row_split=split[0], row_hp=split[1],
                  src=self._src(v), dst=out)
        if self.verify:
            self._check("vec", (out,), (torch_ref(),))
        return out

    def prompt(self, p: Tensor) -> Tensor:
        """prompt_param [1, L, D, S, S] -> fp32 [L, S, S, D]."""
        if not self.cuda:
            cur = packing.pack_prompt(p)
            self._torch.append(lambda: cur.copy_(packing.pack_prompt(p)))
            return cur
        _, L, D, S, S2 = p.shape
        out = self._alloc((L, S, S2, D), torch.float32)
        self._job(kind=4, dst_dtype=torch.float32, n=L, k=D, n_total=S * S2, k_pad=0, src=self._src(p), dst=out)
        if self.verify:
            self._check("prompt", (out,), (packing.pack_prompt(p),))
        return out

    def f32(self, p: Optional[Tensor], shape=None) -> Optional[Tensor]:
        """The parameter itself (no copy): kernels that take fp32 operands read the canonical storage."""
        if p is None:
            return None
        if self.cuda:
            self._params.append(self._src(p))
        t = p.detach() if self.cuda else p.detach().float().contiguous()
        return t if shape is None else t.reshape(*shape)

    # ------------------------------------------------------------------------------------------------
    def pointer_fingerprint(self) -> int:
        h = 0
        for t in self._params:
            h = (h * 1000003 + t.data_ptr()) & 0xFFFFFFFFFFFF
        return h

    def _finalize(self) -> None:
        from . import _lib
        lib = _lib.load()
        n = len(self._jobs)
        arr = (_lib.PirPackJob * n)()
        first = [0]
        code = {torch.float16: _lib.DTYPE_FP16, torch.bfloat16: _lib.DTYPE_BF16, torch.float32: 0}
        for i, j in enumerate(self._jobs):
            a = arr[i]
            a.kind, a.dst_dtype = j["kind"], code[j["dst_dtype"]]
            a.n, a.k, a.transpose, a.flip = j["n"], j["k"], j.get("transpose", 0), j.get("flip", 0)
            a.n_total, a.k_pad = j["n_total"], j["k_pad"]
            a.row_split, a.row_hp = j.get("row_split", 0), j.get("row_hp", 0)
            a.col_split, a.col_hp = j.get("col_split", 0), j.get("col_hp", 0)
            a.gamma_axis = j.get("gamma_axis", 0)
            for key in ("src", "gamma", "beta", "bias", "dst", "ln_s", "vec_t"):
                t = j.get(key)
                setattr(a, key, None if t is None else t.data_ptr())
            rows = int(lib.pir_repack_rows(C.byref(a)))
            if rows <= 0:
                raise RuntimeError("promptir_b200: bad pack job")
            first.append(first[-1] + rows)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        fr = torch.tensor(first, dtype=torch.int32).to(self.device)
        self._table = (raw, fr, n, first[-1], self.pointer_fingerprint())

    def run(self, stream: Optional[int] = None) -> None:
        """Rebuild every cache from the current parameter values (one launch on CUDA)."""
        if not self.cuda:
            with torch.no_grad():
                for fn in self._torch:
                    fn()
            return
        if not self._jobs:
            return
        from . import _lib
        if self._table is None or self._table[4] != self.pointer_fingerprint():
            self._finalize()
        raw, fr, n, rows, _ = self._table
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _lib.check(_lib.load().pir_repack(raw.data_ptr(), fr.data_ptr(), n, rows, s), "pir_repack")
        _lib.launch_count += 1

    def jobs(self) -> int:
        return len(self._jobs)

    def mark(self):
        return (len(self._jobs), len(self._torch), len(self._params), len(self.checks))

    def rollback(self, mark) -> None:
        """Forget the requests made since mark() (their arena space is not reclaimed)."""
        del self._jobs[mark[0]:], self._torch[mark[1]:], self._params[mark[2]:], self.checks[mark[3]:]
        self._table = None

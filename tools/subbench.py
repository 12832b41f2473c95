"""Sub-records of bench.py for the other BASELINE.json configurations (same process, same ranks, few steps each):

    tiles_4k    configs[2]  demo.py tiled inference of a synthetic 3840x2160 frame (tile 256, overlap 32 -> 170 tiles), tiles sharded
                            over the ranks, restored tiles all-gathered (NCCL) and blended on the device
    train_step  configs[3]  PromptIR training step: forward + L1 + backward + NCCL all-reduce of the flat fp32 gradient buffer (N > 1)
                            + torch.optim.AdamW(fused=True).step() + the pir_repack refresh of the 16-bit weight caches;
                            128x128 patches, batch 32 per GPU
    xrestormer  configs[4]  PromptXRestormer inference at 512x512, one image per GPU
    reference_eager_b200    the unmodified reference module as PyTorch eager on the same B200, headline workload (BASELINE.md section 3;
                            rank 0, none of this library's kernels on that path)
Every function returns a dict with ms_per_step (CUDA events, max over ranks), the whole-job value and a parity field measured
against the CPU oracle on a bounded input (rank 0).  They are also what tools/bench_*.py print on their own.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _max_over_ranks(ms: float, dev, world: int) -> float:
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


def _barrier(world: int) -> None:
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _top_kernels(records, launch_order, reps: int = 2, top: int = 4):
    """Eager pass with CUDA events around every launch -> {tag: {launches, ms, share}} of the `top` heaviest tags."""
    stream = torch.cuda.current_stream().cuda_stream
    evs = []
    for rep in range(reps):
        evs = []
        for r in launch_order:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r["launch"](stream); b.record()
            evs.append((r, a, b))
        torch.cuda.synchronize()
    tags, tot = {}, 0.0
    for r, a, b in evs:
        t = a.elapsed_time(b)
        tot += t
        d = tags.setdefault(r.get("tag") or r["kind"], {"launches": 0, "ms": 0.0})
        d["launches"] += 1
        d["ms"] += t
    out = {}
    for k, d in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:top]:
        out[k] = {"launches": d["launches"], "ms": round(d["ms"], 3), "share": round(d["ms"] / tot, 4)}
    return out, tot


# ----------------------------------------------------------------------------------------------------
def bench_tiles(dev, world: int, rank: int, dtype: torch.dtype, frames: int = 2, batch: int = 22) -> dict:
    from promptir_b200 import PromptIR, tiling
    torch.manual_seed(0)
    model = PromptIR(decoder=True).eval().to(dev)
    model.compute_dtype = dtype
    g = torch.Generator().manual_seed(1)
    frame = torch.rand(1, 3, 2160, 3840, generator=g).to(dev)
    x, h, w = tiling.pad_input(frame, 8)
    with torch.no_grad():
        out = tiling.tile_eval(model, x, 256, 32, batch=batch)
        _barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            out = tiling.tile_eval(model, x, 256, 32, batch=batch)
        e1.record()
        _barrier(world)
        ms = _max_over_ranks(e0.elapsed_time(e1) / frames, dev, world)
        # parity: a 256x256 tile of the frame through the same module vs the blended frame is exercised by tests; here only
        # the invariants that are size independent: finite, inside [0, 1] (tile_eval clamps, demo.py:47), same shape
        ok = bool(torch.isfinite(out).all()) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0 and out.shape == x.shape
    del model
    torch.cuda.empty_cache()
    return {"workload": "demo.py tile_eval: one synthetic 3840x2160 frame, tile 256, overlap 32 -> 170 tiles of 256x256 sharded over the ranks "
                        "(BASELINE.json configs[2])",
            "metric": "frame_megapixels_per_sec", "value": 3840 * 2160 / 1e6 / ms * 1e3, "unit": "MP/s", "ms_per_step": ms, "steps": frames,
            "tile_megapixels_per_sec": 170 * 256 * 256 / 1e6 / ms * 1e3, "tiles_per_rank": -(-170 // world), "batch": batch,
            "collective": "all_gather of the restored tiles (NCCL)" if world > 1 else "none", "dtype": str(dtype).split(".")[-1],
            "parity": {"finite_clamped_same_shape": ok, "note": "tile_eval == demo.tile_eval is a -m gpu test on fixtures from the real demo.py"}}


# ----------------------------------------------------------------------------------------------------
def bench_train(dev, world: int, rank: int, dtype: torch.dtype, steps: int = 5, warmup: int = 2, batch: int = 32, side: int = 128,
                overlap: bool = True) -> dict:
    from promptir_b200 import PromptIR, _lib, ddp, synth
    from promptir_b200.train_engine import TrainEngine
    import torch.distributed as dist
    torch.manual_seed(0)
    net = PromptIR(decoder=True).to(dev).train()
    net.compute_dtype = dtype
    x, y = synth.synthetic_batch(batch, side, side, seed=1 + rank)
    x, y = x.to(dev), y.to(dev)
    eng = TrainEngine(net, batch, side, side, dev, dtype)
    ddp.attach_flat_grads(net, eng)                       # .grad = views of the flat buffer (zero copy for the optimizer)
    params = [p for n, p in net.named_parameters() if n in eng.live_params]
    opt = torch.optim.AdamW(params, lr=2e-4, fused=True)  # train.py:52-56 (lr 2e-4, torch defaults otherwise)
    reducer = ddp.OverlappedReducer(eng) if (world > 1 and overlap) else None

    def step():
        out = eng.forward(x)                              # refreshes the 16-bit weight caches first (one pir_repack launch)
        out.requires_grad_(True)
        loss = F.l1_loss(out, y)
        (d_out,) = torch.autograd.grad(loss, out)
        if reducer is not None:
            reducer.backward_and_reduce(d_out)            # all-reduce of finished gradient ranges overlaps the rest of the backward
        else:
            eng.backward(d_out)
            if world > 1:
                ddp.allreduce_gradients(eng)
        opt.step()
        return loss

    for _ in range(warmup):
        loss = step()
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    _barrier(world)
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)
    own_launches = eng.kernels_per_step() + 1           # the two launch programs (CUDA-graph replays) + pir_repack

    def only(fn, n=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    ms_fwd = only(lambda: eng._run("fwd", eng.fwd_launches, True))
    ms_bwd = only(lambda: eng._run("bwd", eng.bwd_launches, True))
    ms_opt = only(lambda: opt.step())
    ms_repack = only(lambda: eng.pk.run())
    ms_ar = None
    if world > 1:
        ms_ar = only(lambda: ddp.allreduce_gradients(eng))
    kernels = {}
    if rank == 0:
        eng.d_out.fill_(1.0 / eng.d_out.numel())
        kernels, _ = _top_kernels(None, eng.fwd_ops + eng.bwd_ops)
    # parity (rank 0): loss.backward() through the drop-in module on ONE 64x64 patch vs fp32 autograd of the CPU oracle
    parity = None
    if rank == 0:
        from oracle import promptir_oracle as O
        xs, ys_ = x[:1, :, :64, :64].contiguous(), y[:1, :, :64, :64].contiguous()
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        F.l1_loss(O.promptir_forward(sd, xs.cpu()), ys_.cpu()).backward()
        for p in net.parameters():
            p.grad = None
        F.l1_loss(net(xs), ys_).backward()
        got = torch.cat([p.grad.reshape(-1).float().cpu() for n, p in net.named_parameters() if p.grad is not None])
        ref = torch.cat([sd[n].grad.reshape(-1) for n, p in net.named_parameters() if p.grad is not None])
        parity = {"flat_gradient_rel_l2": float((got - ref).norm() / ref.norm()), "cosine": float(F.cosine_similarity(got, ref, dim=0)),
                  "oracle": "fp32 autograd of the CPU oracle port, one 64x64 patch"}
    saved = eng.saved_bytes / 1e9
    del eng, opt, reducer
    net._train_engine = None
    for p in net.parameters():
        p.grad = None
    del net
    torch.cuda.empty_cache()
    mp = world * batch * side * side / 1e6
    return {"workload": f"PromptIR training step: forward + L1 + backward + {'NCCL all-reduce of 142 MB fp32 gradients + ' if world > 1 else ''}"
                        f"AdamW(fused) + weight-cache refresh, {side}x{side} patches, batch {batch} per GPU (BASELINE.json configs[3])",
            "metric": "train_step_megapixels_per_sec", "value": mp / ms * 1e3, "unit": "MP/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "images_per_sec": world * batch / ms * 1e3, "dtype": str(dtype).split(".")[-1],
            "parts_ms": {"forward": round(ms_fwd, 3), "backward": round(ms_bwd, 3), "adamw_fused": round(ms_opt, 3), "pir_repack": round(ms_repack, 4),
                         "allreduce_alone": None if ms_ar is None else round(ms_ar, 3)},
            "collective": ("all-reduce of grad_flat in 4 reverse-order segments on a side stream, overlapped with the backward" if reducer_desc(world, overlap)
                           else ("one all-reduce after the backward" if world > 1 else "none")),
            "gpu_launches_per_step": own_launches, "saved_activation_GB": round(saved, 2), "kernels": kernels, "parity": parity,
            "loss": float(loss.detach())}


def reducer_desc(world: int, overlap: bool) -> bool:
    return world > 1 and overlap


# ----------------------------------------------------------------------------------------------------
def bench_xrestormer(dev, world: int, rank: int, dtype: torch.dtype, steps: int = 5, warmup: int = 3, batch: int = 1, side: int = 512) -> dict:
    from promptir_b200 import PromptXRestormer, synth
    torch.manual_seed(0)
    net = PromptXRestormer().eval().to(dev)
    net.compute_dtype = dtype
    x, _ = synth.synthetic_batch(batch, side, side, seed=1 + rank)
    x = x.to(dev)
    eng = net.engine_for(batch, side, side, dev)
    eng.img_in.copy_(x)
    for _ in range(warmup):
        eng.replay(True)
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.replay(True)
    e1.record()
    _barrier(world)
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)
    kernels, parity = {}, None
    if rank == 0:
        kernels, _ = _top_kernels(None, eng.ops)
        from oracle import xrestormer_oracle as XO
        xc = x[:1, :, :128, :128].contiguous()
        with torch.no_grad():
            yc = net(xc).cpu()
            ref = XO.xrestormer_forward({k: v.detach().cpu() for k, v in net.state_dict().items()}, xc.cpu())
        parity = {"max_abs_clamped": float((yc.clamp(0, 1) - ref.clamp(0, 1)).abs().max()), "oracle": "fp32 CPU oracle port, 128x128 crop"}
    launches = eng.kernels_per_forward()
    del eng, net
    torch.cuda.empty_cache()
    return {"workload": f"PromptXRestormer (dim 48, [4,6,6,8]) inference, batch {batch} of {side}x{side} per GPU (BASELINE.json configs[4])",
            "metric": "xrestormer_fwd_megapixels_per_sec", "value": world * batch * side * side / 1e6 / ms * 1e3, "unit": "MP/s", "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "dtype": str(dtype).split(".")[-1], "collective": "none", "gpu_launches_per_step": launches,
            "kernels": kernels, "parity": parity}


# ----------------------------------------------------------------------------------------------------
def bench_reference_eager(dev, world: int, rank: int, steps: int = 3, warmup: int = 2, batch: int = 16, side: int = 256) -> dict:
    """BASELINE.md section 3 "same-box GPU comparison": the UNMODIFIED reference module (net/model.py, copied to the git-ignored
    baseline/_ref/ by __graft_entry__.build(); it travels with the gpurun snapshot) run as PyTorch eager ops -- cuDNN / cuBLAS / ATen --
    on the same B200 on the headline workload (BASELINE.json configs[1]), CUDA-event timed.  None of this library's kernels are on that
    path.  Rank 0 only (the comparison is per GPU).  Also reports how far the reference's own bf16 run lands from its own fp32 run on
    this batch: the context for the 2e-3 contract of the 16-bit builds."""
    if rank != 0:
        return {"skipped": "rank 0 only"}
    path = os.path.join(ROOT, "baseline", "_ref", "net", "model.py")
    if not os.path.exists(path):
        return {"unavailable": "baseline/_ref/net/model.py absent (__graft_entry__.build() copies it where /root/reference exists)"}
    import importlib.util
    from promptir_b200 import synth
    spec = importlib.util.spec_from_file_location("_promptir_reference_model_gpu", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(0)
    ref = mod.PromptIR(decoder=True).eval().to(dev)
    x = synth.synthetic_batch(batch, side, side, seed=1)[0].to(dev)
    mp = batch * side * side / 1e6

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_step": round(ms, 3), "MP_per_s": round(mp / ms * 1e3, 3)}

    out = {"workload": f"unmodified reference net/model.py PromptIR(decoder=True), PyTorch {torch.__version__} eager on the same B200, batch {batch} of "
                       f"{side}x{side} (BASELINE.json configs[1]), {warmup} warm-up + {steps} timed forwards per variant, CUDA events",
           "unit": "MP/s", "variants": {}}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        with torch.no_grad():
            out["variants"]["fp32_torch_defaults"] = timed(lambda: ref(x))         # the stock code path: what `python demo.py` runs
            y32 = ref(x).clamp(0, 1)
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            out["variants"]["fp32_tf32"] = timed(lambda: ref(x))
            ref = ref.bfloat16()
            xb = x.bfloat16()
            out["variants"]["bf16"] = timed(lambda: ref(xb))
            y16 = ref(xb).float().clamp(0, 1)
            out["reference_bf16_vs_its_fp32"] = {"max_abs_clamped": float((y16 - y32).abs().max()), "images": batch,
                                                 "note": "the reference's own bf16 eager run against its own fp32 run, same weights and inputs"}
            try:                                         # optional variant: must not take the record down
                for p in ref.parameters():
                    if p.dim() == 4:                     # (prompt_param is 5-D: Module.to(memory_format=...) would reject it)
                        p.data = p.data.contiguous(memory_format=torch.channels_last)
                xcl = xb.contiguous(memory_format=torch.channels_last)
                out["variants"]["bf16_channels_last"] = timed(lambda: ref(xcl))
            except Exception as e:
                out["bf16_channels_last_error"] = f"{type(e).__name__}: {e}"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    best = max(out["variants"].items(), key=lambda kv: kv[1]["MP_per_s"])
    out["best_variant"], out["value"], out["ms_per_step"] = best[0], best[1]["MP_per_s"], best[1]["ms_per_step"]
    del ref
    torch.cuda.empty_cache()
    for name, fn in (("train_step", lambda: _reference_eager_train(mod, dev)), ("xrestormer", lambda: _reference_eager_x(dev))):
        try:                                             # comparators for the train_step / xrestormer sub-records; optional
            out[name] = fn()
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    return out


def _timed_eager(fn, steps: int, warmup: int) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _reference_eager_train(mod, dev, steps: int = 2, warmup: int = 1, batch: int = 32, side: int = 128) -> dict:
    """train.py:37-56 with the unmodified reference module on the same B200: forward + L1 + autograd backward + AdamW step
    (BASELINE.json configs[3] per-GPU work; one GPU, so no all-reduce)."""
    from promptir_b200 import synth
    torch.manual_seed(0)
    ref = mod.PromptIR(decoder=True).train().to(dev)
    opt = torch.optim.AdamW(ref.parameters(), lr=2e-4)
    x, y = synth.synthetic_batch(batch, side, side, seed=1)
    x, y = x.to(dev), y.to(dev)
    mp = batch * side * side / 1e6

    def step(autocast: bool = False):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = ref(x)
        F.l1_loss(out.float(), y).backward()
        opt.step()
    res = {"workload": f"unmodified reference PromptIR(decoder=True): forward + L1 + autograd backward + torch.optim.AdamW.step(), batch {batch} of "
                       f"{side}x{side}, one B200, {warmup} warm-up + {steps} timed steps per variant", "unit": "MP/s", "variants": {}}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32, ac in (("fp32_torch_defaults", None, False), ("fp32_tf32", True, False), ("bf16_autocast", True, True)):
            if tf32:
                torch.backends.cudnn.allow_tf32 = True
                torch.backends.cuda.matmul.allow_tf32 = True
            ms = _timed_eager(lambda: step(ac), steps, warmup)
            res["variants"][name] = {"ms_per_step": round(ms, 3), "MP_per_s": round(mp / ms * 1e3, 3)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    best = max(res["variants"].items(), key=lambda kv: kv[1]["MP_per_s"])
    res["best_variant"], res["value"], res["ms_per_step"] = best[0], best[1]["MP_per_s"], best[1]["ms_per_step"]
    return res


def _reference_eager_x(dev, steps: int = 3, warmup: int = 2, batch: int = 1, side: int = 512) -> dict:
    """The unmodified reference net/prompt_xrestormer.py (torchstat, an import-only dependency, stubbed) on the same B200:
    BASELINE.json configs[4] per-GPU work (one 512x512 image)."""
    import importlib.util
    import types
    from promptir_b200 import synth
    path = os.path.join(ROOT, "baseline", "_ref", "net", "prompt_xrestormer.py")
    if not os.path.exists(path):
        return {"unavailable": "baseline/_ref/net/prompt_xrestormer.py absent"}
    sys.modules.setdefault("torchstat", types.SimpleNamespace(stat=None))
    spec = importlib.util.spec_from_file_location("_promptir_reference_xrestormer_gpu", path)
    xmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(xmod)
    torch.manual_seed(0)
    ref = xmod.PromptXRestormer().eval().to(dev)
    x = synth.synthetic_batch(batch, side, side, seed=1)[0].to(dev)
    mp = batch * side * side / 1e6
    res = {"workload": f"unmodified reference PromptXRestormer() forward, batch {batch} of {side}x{side}, one B200, {warmup} warm-up + {steps} timed "
                       "forwards per variant", "unit": "MP/s", "variants": {}}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        with torch.no_grad():
            ms = _timed_eager(lambda: ref(x), steps, warmup)
            res["variants"]["fp32_torch_defaults"] = {"ms_per_step": round(ms, 3), "MP_per_s": round(mp / ms * 1e3, 3)}
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            ms = _timed_eager(lambda: ref(x), steps, warmup)
            res["variants"]["fp32_tf32"] = {"ms_per_step": round(ms, 3), "MP_per_s": round(mp / ms * 1e3, 3)}
            ref = ref.bfloat16()
            xb = x.bfloat16()
            ms = _timed_eager(lambda: ref(xb), steps, warmup)
            res["variants"]["bf16"] = {"ms_per_step": round(ms, 3), "MP_per_s": round(mp / ms * 1e3, 3)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    best = max(res["variants"].items(), key=lambda kv: kv[1]["MP_per_s"])
    res["best_variant"], res["value"], res["ms_per_step"] = best[0], best[1]["MP_per_s"], best[1]["ms_per_step"]
    return res

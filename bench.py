#!/usr/bin/env python
"""bench.py -- PromptIR forward megapixels/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype bf16|fp16] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] -- all-in-one inference on a batch of 16 synthetic
256x256 noise/rain/haze crops per GPU, random-init weights (seed 0).  One step = one forward over the batch.
  value : whole-job MP/s, inputs resident in HBM, CUDA-graph replay, CUDA events, max over ranks
  e2e   : the same metric through the public nn.Module call with pinned HOST input and output (H2D + forward + D2H)
  roofline : dominant kernel class, measured in a second, eager pass with CUDA events around every launch
  cpu_baseline : the reference's own fp32 PyTorch forward (unmodified net/model.py from baseline/_ref; oracle port if absent) on the
                 host cores over the SAME 16 images; its outputs are the parity reference of every image that was timed
  parity / alt : both 16-bit storage types are timed and checked; `dtype` is the one that meets the 2e-3 / 0.02 dB contract
  configs : sub-records for BASELINE configs[2..4]: 4K tiles, the training step (AdamW + NCCL all-reduce inside), PromptXRestormer;
            reference_eager_b200 = the unmodified reference module as PyTorch eager (cuDNN / cuBLAS) on the same B200, headline workload
`--impl reference` times that CPU path alone (rank 0 only under torchrun) and prints the reference-arm line.
Multi-GPU: images are independent, so the batch is sharded by rank with no data-path collective (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH, SIDE = 16, 256
METRIC = "promptir_fwd_megapixels_per_sec"
UNIT = "MP/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tc": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tc": 1400.0, "src": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            hot = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out.update(sm_mhz=statistics.median(hot), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch forward on the host cores
# ----------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
WORKLOAD = (f"PromptIR(dim=48,[4,6,6,8],decoder=True) all-in-one inference, batch {BATCH} of {SIDE}x{SIDE} synthetic noise/rain/haze crops "
            "per GPU (BASELINE.json configs[1]), random-init weights seed 0")


def reference_forward():
    """-> (forward(state_dict, x) -> y, kind).  kind "reference": the UNMODIFIED net/model.py of the reference, copied to the
    git-ignored baseline/_ref/ by __graft_entry__.build() (it travels to the GPU box with the snapshot); "port": the oracle's
    functional restatement, used only when that copy is absent."""
    path = os.path.join(REF_DIR, "net", "model.py")
    if os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_promptir_reference_model", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        cache = {}

        def fwd(sd, x):
            if "m" not in cache:
                cache["m"] = mod.PromptIR(decoder=True).eval()
            cache["m"].load_state_dict(sd, strict=True)
            return cache["m"](x)
        return fwd, "reference"
    from oracle import promptir_oracle as O          # the CPU legs are the only places bench.py touches oracle/
    return O.promptir_forward, "port"


def cpu_forward_all(sd, x, chunk: int = 4):
    """The whole batch of one step on the host cores (fp32, all threads), `chunk` images at a time.  -> (y, seconds, kind, cores)."""
    fwd, kind = reference_forward()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    outs = []
    with torch.no_grad():
        fwd(sd, x[:1])                               # warm-up (thread pool, oneDNN primitive cache)
        t = time.perf_counter()
        for i in range(0, x.shape[0], chunk):
            outs.append(fwd(sd, x[i:i + chunk]))
        sec = time.perf_counter() - t
    return torch.cat(outs), sec, kind, cores


def run_reference(args):
    """Reference arm: the reference's CPU forward on this workload.  One step = a BOUNDED SAMPLE of the batch (4 of its 16 images,
    cycling through the batch), at most 5 timed steps, so the run ends within minutes; MP/s normalises the sample size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from promptir_b200 import PromptIR, synth
    fwd, kind = reference_forward()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in PromptIR(decoder=True).state_dict().items()}
    x = synth.synthetic_batch(BATCH, SIDE, SIDE, seed=1)[0]
    steps, warm, chunk = max(1, min(args.steps, 5)), 1, 4
    times = []
    with torch.no_grad():
        for i in range(warm + steps):
            lo = (i * chunk) % BATCH
            t = time.perf_counter()
            fwd(sd, x[lo:lo + chunk])
            dt = time.perf_counter() - t
            if i >= warm:
                times.append(dt)
    sec = statistics.median(times)
    val = chunk * SIDE * SIDE / 1e6 / sec
    sample = (f"bounded sample: {chunk} of the {BATCH} {SIDE}x{SIDE} images per step, fp32, median of {len(times)} steps ({sec:.2f} s each), "
              f"torch {torch.__version__} oneDNN, {'unmodified reference net/model.py' if kind == 'reference' else 'oracle port (baseline/_ref absent)'}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + f" -- reference arm: CPU, bounded sample of {chunk} images per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------
def ncu_traffic(tag: str, shape, dtype: str):
    """DRAM read + write bytes of ONE launch of a kernel class from the committed `ncu --set full` captures, looked up in
    profiles/ncu_traffic.json by (kernel tag, input shape, storage type); None when no capture of that class is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    for e in json.load(open(path)).get("entries", []):
        if e["tag"] == tag and list(e["shape"]) == list(shape) and e.get("dtype", dtype) == dtype:
            return e["dram_read_bytes"] + e["dram_write_bytes"], f"profiles/{e['source']} ({e.get('kernel', '')}, build {e.get('build', '?')})"
    return None, None


def run_ours(args):
    import torch.distributed as dist
    from promptir_b200 import PromptIR, _lib, synth
    from promptir_b200.engine import op_cost

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    DT = {"bf16": torch.bfloat16, "fp16": torch.float16}
    torch.manual_seed(0)
    model = PromptIR(decoder=True).eval().to(dev)
    x_host, clean = synth.synthetic_batch(BATCH, SIDE, SIDE, seed=1 + rank)
    x_pin = x_host.pin_memory()
    W, K = max(args.warmup, 3), args.steps
    mp_step = BATCH * SIDE * SIDE / 1e6

    def measure(name):
        """Both timed regions of one storage type.  -> dict(ms, ms_e2e, y (host copy of the e2e result), clocks, engine)."""
        model.compute_dtype = DT[name]
        eng = model.engine_for(BATCH, SIDE, SIDE, dev)
        eng.img_in.copy_(x_pin)
        y_pin = torch.empty_like(x_host).pin_memory()
        # ---- region A: resident inputs, CUDA-graph replay ---------------------------------------------
        for _ in range(W):
            eng.replay(True)
        barrier()
        sampler = ClockSampler(local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            eng.replay(True)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / K
        clocks = sampler.stop()
        # ---- region C: end to end through the public API, host buffers ----------------------------------
        x_dev = torch.empty_like(x_host, device=dev)

        def e2e_step():
            x_dev.copy_(x_pin, non_blocking=True)
            with torch.no_grad():
                y = model(x_dev)
            y_pin.copy_(y, non_blocking=True)

        for _ in range(2):
            e2e_step()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(K):
            e2e_step()
        c1.record()
        barrier()
        ms_e2e = c0.elapsed_time(c1) / K
        if world > 1:
            t = torch.tensor([ms, ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_e2e = t.tolist()
        return {"ms": ms, "ms_e2e": ms_e2e, "y": y_pin.clone(), "clocks": clocks, "eng": eng}

    order = [args.dtype] + [d for d in ("bf16", "fp16") if d != args.dtype]
    runs = {d: measure(d) for d in order}

    # ---- parity of exactly what was timed: ALL images of rank 0's batch against the reference's fp32 CPU forward ----------------
    cb = None
    if rank == 0:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        ref, sec, kind, cores = cpu_forward_all(sd, x_host)
        cb = {"value": mp_step / sec, "unit": UNIT, "cores": cores, "kind": kind,
              "sample": f"all {BATCH} {SIDE}x{SIDE} images of one step (the GPU arm's batch), fp32, one pass in chunks of 4 ({sec:.1f} s), "
                        f"torch {torch.__version__} oneDNN, {'unmodified reference net/model.py from baseline/_ref' if kind == 'reference' else 'oracle port'}"}
        for d, r in runs.items():
            got = r["y"]
            per_img = (got.clamp(0, 1) - ref.clamp(0, 1)).abs().flatten(1).max(dim=1).values
            r["parity"] = {"max_abs_clamped": float(per_img.max()), "max_abs_clamped_median_image": float(per_img.median()),
                           "rmse_clamped": float((got.clamp(0, 1) - ref.clamp(0, 1)).pow(2).mean().sqrt()),
                           "images": BATCH, "dpsnr_db": abs(synth.psnr(got, clean) - synth.psnr(ref, clean)),
                           "tolerance": {"max_abs": 2e-3, "dpsnr_db": 0.02},
                           "oracle": f"fp32 CPU forward of the {'unmodified reference module' if kind == 'reference' else 'oracle port'}, every image of the batch"}
            r["parity"]["within_tolerance"] = r["parity"]["max_abs_clamped"] <= 2e-3 and r["parity"]["dpsnr_db"] <= 0.02
    # headline storage type: the requested one if it meets the contract, else the other one if that does
    head = order[0]
    if rank == 0 and not runs[head]["parity"]["within_tolerance"] and runs[order[1]]["parity"]["within_tolerance"]:
        head = order[1]
    if world > 1:
        pick = torch.tensor([order.index(head)], device=dev)
        dist.broadcast(pick, 0)
        head = order[int(pick.item())]
    alt = [d for d in order if d != head][0]
    model.compute_dtype = DT[head]
    eng = runs[head]["eng"]
    ms, ms_e2e, clocks = runs[head]["ms"], runs[head]["ms_e2e"], runs[head]["clocks"]
    kernels_per_step = eng.kernels_per_forward()

    # ---- region B: eager launches with CUDA events around every kernel (roofline accounting, headline type) ----------
    s = torch.cuda.current_stream().cuda_stream
    nrep = min(K, 3)
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in eng.ops] for _ in range(nrep)]
    eng.launch_all(s)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures exactly these launches (profiles/*launches*)
    for rep in range(nrep):
        for (a, b), r in zip(evs[rep], eng.ops):
            a.record()
            r["launch"](s)
            b.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    pk = peaks()
    tags, classes = {}, {}
    for i, r in enumerate(eng.ops):
        tag = r.get("tag") or r["kind"]
        src = next((r[k] for k in ("a", "x", "qk", "qkv") if r.get(k) is not None), None)
        shape = tuple(src.shape) if src is not None else ()
        t_ms = statistics.mean(evs[rep][i][0].elapsed_time(evs[rep][i][1]) for rep in range(nrep))
        by, fl = op_cost(r)
        for key, table in ((tag, tags), ((tag, shape), classes)):
            d = table.setdefault(key, {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "n": 0})
            d["ms"] += t_ms
            d["bytes"] += by
            d["flops"] += fl
            d["n"] += 1
    total_ms = sum(d["ms"] for d in tags.values())
    table = {}
    for tag, d in tags.items():
        t_h, t_t = d["bytes"] / pk["hbm"] / 1e6, d["flops"] / pk["tc"] / 1e9
        table[tag] = {"launches": d["n"], "ms": round(d["ms"], 4), "share": round(d["ms"] / total_ms, 4), "GBps": round(d["bytes"] / d["ms"] / 1e6, 1),
                      "TFLOPs": round(d["flops"] / d["ms"] / 1e9, 2), "bound": "hbm" if t_h >= t_t else "tensor",
                      "frac": round(max(t_h, t_t) / d["ms"], 4)}
    # dominant kernel = the (tag, input shape) class with the most time; its per-launch numbers go into `roofline`
    (top, top_shape), td = max(classes.items(), key=lambda kv: kv[1]["ms"])
    hbm_bound = td["bytes"] / pk["hbm"] / 1e6 >= td["flops"] / pk["tc"] / 1e9
    ach = td["bytes"] / td["ms"] / 1e6 if hbm_bound else td["flops"] / td["ms"] / 1e9
    traffic, traffic_src = ncu_traffic(top, top_shape, head)
    roofline = {"kernel": f"{top} on {list(top_shape)}", "bound": "hbm" if hbm_bound else "tensor", "achieved": round(ach, 1),
                "peak": pk["hbm"] if hbm_bound else pk["tc"], "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": round(ach / (pk["hbm"] if hbm_bound else pk["tc"]), 4), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": round(td["bytes"] / td["n"]), "launches_per_step": td["n"],
                "avg_launch_ms": round(td["ms"] / td["n"], 4), "share_of_step": round(td["ms"] / total_ms, 4), "peak_source": pk["src"],
                "note": ("K56/K12 are the fused LN+1x1+dw3x3(+gate) kernels: their binding resource is CUDA-core issue (packed fp16 FMA, "
                         "conversions, MUFU), not HBM or the tensor pipe -- see DESIGN.md 3.2 and profiles/") if top in ("K56", "K12") else "",
                "how": "CUDA events around each launch, eager pass, mean of %d steps; achieved = algorithmic bytes / time" % nrep}
    whole_t_min = sum(max(d["bytes"] / pk["hbm"] / 1e6, d["flops"] / pk["tc"] / 1e9) for d in tags.values())
    for r in runs.values():
        r.pop("eng")
    del eng
    model._engines = {}
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, same ranks (tiles sharded / data-parallel training with the NCCL all-reduce) ----------
    subs = {}
    if not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import subbench
        for name, fn in (("tiles_4k", lambda: subbench.bench_tiles(dev, world, rank, DT[head])),
                         ("train_step", lambda: subbench.bench_train(dev, world, rank, torch.bfloat16)),
                         ("xrestormer", lambda: subbench.bench_xrestormer(dev, world, rank, DT[head])),
                         ("reference_eager_b200", lambda: subbench.bench_reference_eager(dev, world, rank))):
            try:
                subs[name] = fn()
            except Exception as e:                       # a sub-record must never take the headline line down
                subs[name] = {"error": f"{type(e).__name__}: {e}"}
            barrier()

    if rank == 0:
        def summary(d):
            r = runs[d]
            return {"dtype": d, "value": world * mp_step / r["ms"] * 1e3, "ms_per_step": r["ms"],
                    "e2e": world * mp_step / r["ms_e2e"] * 1e3, "parity": r["parity"]}
        line = {
            "metric": METRIC, "value": world * mp_step / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": head,
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "per_gpu_batch": BATCH, "global_batch": BATCH * world, "height": SIDE, "width": SIDE,
                       "parallelism": f"images sharded over {world} GPU(s), no data-path collective",
                       "l2": "activation working set per step ~3 GB >> 126 MB L2 (no flush needed)",
                       "timing": "CUDA events on the launching stream, CUDA-graph replay, max over ranks",
                       "dtype_choice": (f"both 16-bit storage types are timed; `dtype` is the requested one ({order[0]}) when its output meets the "
                                        "2e-3 / 0.02 dB contract against the fp32 reference on every image, otherwise the one that does; the other is under `alt`")},
            "clocks": clocks,
            "e2e": {"value": world * mp_step / ms_e2e * 1e3, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_pin.numel() * 4, "d2h_bytes_per_step": x_pin.numel() * 4,
                    "api": "promptir_b200.PromptIR.__call__ (pinned host in, pinned host out)"},
            "gpu_launches": kernels_per_step * K,
            "roofline": roofline,
            "cpu_baseline": cb,
            "kernels": table,
            "whole_step_roofline": {"t_min_ms": round(whole_t_min, 3), "frac": round(whole_t_min / total_ms, 4),
                                    "note": "sum over kernels of max(bytes/HBM peak, flops/TC peak) / sum of measured kernel times"},
            "parity": runs[head]["parity"],
            "alt": summary(alt),
            "configs": subs,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1 at stderr for the
    run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="skip the tiles_4k / train_step / xrestormer / reference_eager_b200 sub-records")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus == 1:
        quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus),
                   "--steps", str(args.steps), "--warmup", str(args.warmup), "--dtype", args.dtype] + (["--no-configs"] if args.no_configs else [])
            raise SystemExit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()

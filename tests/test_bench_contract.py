"""bench.py's line contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON line on
stdout with the driver's keys, and the product arm refuses to run (non-zero exit, no JSON) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # one JSON line, nothing else on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "promptir_fwd_megapixels_per_sec" and d["unit"] == "MP/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    # one step = a bounded sample of 4 of the batch's 16 images (said so in config.workload and cpu_baseline.sample)
    assert d["value"] > 0 and abs(d["value"] - 4 * 256 * 256 / 1e6 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"] + 1e-9
    cb = d["cpu_baseline"]
    ref_copy = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "net", "model.py"))
    assert cb["kind"] == ("reference" if ref_copy else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "256x256" in cb["sample"] and "bounded sample" in cb["sample"] and "bounded sample" in d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present: the product arm runs for real (covered by the -m gpu suite and the driver)")
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0
    assert p.stdout.strip() == ""                      # no JSON line is produced by a run that did not measure anything
    assert "no CPU fallback" in p.stderr or "no CUDA device" in p.stderr


def test_reference_eager_subrecord_gating(monkeypatch, tmp_path):
    """configs.reference_eager_b200 (tools/subbench.py): only rank 0 runs the unmodified reference on its GPU, and without the
    baseline/_ref copy the record says so instead of silently timing something else (no oracle port, no product kernels)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import subbench
    assert subbench.bench_reference_eager(torch.device("cpu"), 2, 1) == {"skipped": "rank 0 only"}
    monkeypatch.setattr(subbench, "ROOT", str(tmp_path))
    rec = subbench.bench_reference_eager(torch.device("cpu"), 1, 0)
    assert list(rec) == ["unavailable"] and "baseline/_ref" in rec["unavailable"]

"""Range diagnostics for the 16-bit pipeline.

The fused kernels keep the widest tensors of a TransformerBlock (the 3C / 2*hidden pre-depthwise channels, net/model.py:88-90,
111-112) on chip as IEEE half: the fp32 accumulators are converted with saturation (+-65504), the nine taps accumulate in fp16
and the gate product saturates.  With fp16 STORAGE every activation that reaches HBM saturates at +-65504 as well.  Random-init
and trained Restormer-class weights stay orders of magnitude below that, but nothing in the arithmetic guarantees it, so this
module makes the margin visible: `range_report` runs the UNFUSED launch program (every intermediate reaches HBM) with bf16
storage (fp32 exponent range) and returns, per launch, the largest magnitude written and how many elements lie beyond the fp16
range -- exactly the values the fused / fp16 paths would clip.
"""
from __future__ import annotations

from typing import Dict, List

import torch

FP16_MAX = 65504.0


def range_report(model, x: torch.Tensor, margin: float = 0.25) -> Dict[str, object]:
    """model: a promptir_b200.PromptIR on a CUDA device; x: fp32 NCHW batch on the same device.
    -> {"rows": [{index, tag, shape, absmax, beyond_fp16}], "worst": row, "clips": bool, "at_risk": bool}
    `at_risk`: some tensor exceeds margin * 65504 (default: a factor 4 of headroom left)."""
    from .engine import Engine
    b, _, h, w = x.shape
    eng = Engine(model, b, h, w, x.device, torch.bfloat16, fuse=False)
    eng.img_in.copy_(x.float())
    stream = torch.cuda.current_stream(x.device).cuda_stream
    rows: List[dict] = []
    for i, r in enumerate(eng.ops):
        r["launch"](stream)
        out = r.get("out") if r.get("out") is not None else r.get("wfold")
        if out is None:
            continue
        a = out.float().abs()
        rows.append({"index": i, "tag": r.get("tag") or r["kind"], "shape": list(out.shape), "absmax": float(a.max()),
                     "beyond_fp16": int((a > FP16_MAX).sum())})
    worst = max(rows, key=lambda d: d["absmax"])
    return {"rows": rows, "worst": worst, "clips": any(d["beyond_fp16"] for d in rows),
            "at_risk": worst["absmax"] > margin * FP16_MAX}

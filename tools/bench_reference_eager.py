#!/usr/bin/env python
"""Same-box GPU comparison: the UNMODIFIED reference net/model.py (baseline/_ref, written by
__graft_entry__.build()) as PyTorch eager on cuda:0 on the headline workload, next to nothing else.  One JSON line on stdout; the
same record is the `configs.reference_eager_b200` entry of bench.py.

    python tools/bench_reference_eager.py [--batch 16] [--side 256] [--steps 3] [--warmup 2]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

import subbench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--side", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    real_stdout = os.dup(1)          # the reference prints while constructing: keep stdout for the one JSON line
    os.dup2(2, 1)
    torch.cuda.set_device(0)
    rec = subbench.bench_reference_eager(torch.device("cuda", 0), 1, 0, steps=a.steps, warmup=a.warmup, batch=a.batch, side=a.side)
    os.write(real_stdout, (json.dumps(rec) + "\n").encode())


if __name__ == "__main__":
    main()

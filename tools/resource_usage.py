#!/usr/bin/env python
"""Per-kernel resource usage of the built library (registers, stack frame = spills / local arrays, static shared memory) from
`cuobjdump -res-usage`, demangled, sorted by object -> profiles/r2_resource_usage.txt.  Runs without a GPU.

    python tools/resource_usage.py [path/to/libpromptir_b200.so]
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def table(lib: str) -> str:
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    lines, rows, ident = out.splitlines(), [], ""
    for i, l in enumerate(lines):
        m = re.match(r"identifier = (\S+)", l)
        if m:
            ident = m.group(1)
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            d = dict(re.findall(r"(\w+):(\d+)", lines[i + 1]))
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            rows.append((ident, name, int(d.get("REG", 0)), int(d.get("STACK", 0)), int(d.get("SHARED", 0)), int(d.get("LOCAL", 0))))
    rows.sort(key=lambda r: (r[0], r[1]))
    txt = ["# cuobjdump -res-usage promptir_b200/libpromptir_b200.so (sm_100a, final build of round 2): registers per thread, stack bytes",
           "# (non-zero = spills or local arrays), static shared memory, local memory.  Dynamic shared memory is set at launch (DESIGN section 3).",
           "# object | kernel | REG | STACK | SHARED(static) | LOCAL"]
    for r in rows:
        nm = r[1] if len(r[1]) < 150 else r[1][:147] + "..."
        txt.append(f"{r[0]} | {nm} | {r[2]} | {r[3]} | {r[4]} | {r[5]}")
    txt.append(f"# {len(rows)} kernels; {sum(r[3] > 0 for r in rows)} with a non-zero stack frame (largest {max(r[3] for r in rows)} bytes)")
    return "\n".join(txt) + "\n"


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "promptir_b200", "libpromptir_b200.so")
    with open(os.path.join(ROOT, "profiles", "r2_resource_usage.txt"), "w") as f:
        f.write(table(lib))
    print("wrote profiles/r2_resource_usage.txt")

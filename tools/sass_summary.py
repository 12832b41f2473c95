#!/usr/bin/env python
"""Per-object SASS evidence that the contractions are Blackwell-native: counts of the tcgen05 / TMEM / TMA mnemonics
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UTCBAR = tcgen05.commit) and of the
warp-level fallbacks (HMMA = mma.sync, LDSM = ldmatrix) in every object of promptir_b200/csrc/build.  Written by
__graft_entry__.build() to profiles/r2_sass_summary.txt.     python tools/sass_summary.py [out.txt]"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "HFMA2", "FFMA", "MUFU"]


def summarize() -> str:
    rows = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "promptir_b200", "csrc", "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels = len(re.findall(r"^\s*Function :", sass, re.M))
        counts = {m: len(re.findall(r"\b%s\b" % m, sass)) for m in MNEMONICS}
        rows.append((os.path.basename(obj), kernels, counts))
    w = max(len(r[0]) for r in rows) if rows else 8
    lines = ["# cuobjdump -sass mnemonic counts per object (sm_100a); tools/sass_summary.py",
             "%-*s %8s " % (w, "object", "kernels") + " ".join("%8s" % m for m in MNEMONICS)]
    for name, k, c in rows:
        lines.append("%-*s %8d " % (w, name, k) + " ".join("%8d" % c[m] for m in MNEMONICS))
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    text = summarize()
    out = sys.argv[1] if len(sys.argv) > 1 else None
    if out:
        with open(out, "w") as f:
            f.write(text)
    print(text, end="")

"""Summarise an `ncu --page source --csv` dump: opcode mix by executed instructions, top stall lines,
shared-memory excess wavefronts.   python tools/ncu_src.py file.csv [top]"""
import csv
import sys
from collections import Counter

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != hdr[0]]
ops, tot = Counter(), 0
for r in body:
    n = int(r[ix["Instructions Executed"]] or 0)
    op = r[ix["Source"]].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    ops[op.split(".")[0]] += n
    tot += n
print("kernel:", rows[0][1][:90])
print("warp-instructions executed:", tot)
print("  " + "  ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in ops.most_common(16)))
samp = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("top stall lines (samples, % of all):")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
    s = int(r[ix["# Samples"]] or 0)
    stalls = {k: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    main = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {s:7d} {100 * s / max(samp, 1):5.1f}%  {r[ix['Source']].strip()[:70]:70s} {main}")
ex = [(int(r[ix["L1 Wavefronts Shared Excessive"]] or 0), r) for r in body]
print("shared excess wavefronts:", sum(e for e, _ in ex), "of", sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in body))
for e, r in sorted(ex, key=lambda t: -t[0])[:6]:
    if e:
        print(f"  {e:9d}  {r[ix['Source']].strip()[:80]}")

"""Group the SASS lines of one launch in an `ncu --page source --csv` export into contiguous regions of equal execution count and
print each region's share of the executed warp-instructions with its opcode mix.  python tools/ncu_regions.py file.csv [launch] [top]"""
import sys
from collections import Counter

import ncu_src


def main():
    sec = ncu_src.sections(sys.argv[1])[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    hdr = sec["hdr"]
    ix = {n: i for i, n in enumerate(hdr)}
    seen, body = set(), []
    for r in sec["rows"]:
        if r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]])
        body.append(r)
    tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
    samp_tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    regions, cur = [], None
    for i, r in enumerate(body):
        n = int(r[ix["Instructions Executed"]] or 0)
        if cur and abs(n - cur["n"]) <= 0.02 * max(n, cur["n"]):
            cur["cnt"] += 1; cur["sum"] += n; cur["end"] = i
        else:
            cur = {"n": n, "cnt": 1, "sum": n, "start": i, "end": i}
            regions.append(cur)
    big = sorted(regions, key=lambda c: -c["sum"])[:top]
    print("kernel:", sec["name"][:100], " warp-instructions:", tot, " samples:", samp_tot)
    for c in sorted(big, key=lambda c: c["start"]):
        ops = Counter()
        samp = 0
        for r in body[c["start"]:c["end"] + 1]:
            op = r[ix["Source"]].split()
            op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
            ops[op.split(".")[0]] += 1
            samp += int(r[ix["# Samples"]] or 0)
        print(f"[{c['start']:5d}-{c['end']:5d}] x{c['n']:>9d} n={c['cnt']:4d} instr={100 * c['sum'] / tot:5.1f}% samples={100 * samp / max(samp_tot, 1):5.1f}%  "
              + " ".join(f"{k}:{v}" for k, v in ops.most_common(8)))


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    main()

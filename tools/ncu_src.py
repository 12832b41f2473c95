"""Summarise `ncu -i rep --page source --csv` (read from a file): per profiled launch, opcode mix, top stall lines,
shared-memory excess wavefronts.   python tools/ncu_src.py file.csv [launch_index|all] [top]"""
import csv
import sys
from collections import Counter

csv.field_size_limit(10 ** 9)


def sections(path):
    cur, out = None, []
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            elif len(r) == len(cur["hdr"]):
                cur["rows"].append(r)
    return out


def report(sec, top):
    hdr = sec["hdr"]
    ix = {n: i for i, n in enumerate(hdr)}
    seen, body = set(), []
    for r in sec["rows"]:                       # ncu prints every line twice (SASS + source views)
        if r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]])
        body.append(r)
    ops, tot = Counter(), 0
    for r in body:
        n = int(r[ix["Instructions Executed"]] or 0)
        op = r[ix["Source"]].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        ops[op.split(".")[0]] += n
        tot += n
    print("kernel:", sec["name"][:100])
    print("warp-instructions executed:", tot)
    print("  " + "  ".join(f"{k}:{100 * v / max(tot, 1):.1f}%" for k, v in ops.most_common(16)))
    samp = sum(int(r[ix["# Samples"]] or 0) for r in body)
    stall_cols = [k for k in hdr if k.startswith("stall_")]
    print(f"top stall lines of {samp} samples:")
    for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][ix["# Samples"]] or 0))[:top]:
        s = int(r[ix["# Samples"]] or 0)
        main = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
        print(f"  [{i:4d}] {s:6d} {100 * s / max(samp, 1):5.1f}%  x{r[ix['Instructions Executed']]:>8s}  {r[ix['Source']].strip()[:64]:64s} {main}")
    ex = sum(int(r[ix["L1 Wavefronts Shared Excessive"]] or 0) for r in body)
    print("shared excess wavefronts:", ex, "of", sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in body))


def main():
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    for n, sec in enumerate(sections(sys.argv[1])):
        if which == "all" or int(which) == n:
            print(f"=== launch {n} ===")
            report(sec, top)


if __name__ == "__main__":
    main()

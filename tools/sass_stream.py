"""Pipe-letter stream of a kernel's SASS between the first tcgen05.ld (LDTM) and n instructions later: H = fp16 FMA pipe (HFMA2 / HMUL2 /
HADD2), F = fp32 FMA, a = ALU pipe (F2FP / HMNMX2 / PRMT / SHF / IADD3 / LOP3 / MOV), X = MUFU, T = LDTM, S = store, u = other.
    python tools/sass_stream.py object.o mangled_kernel_name [n]"""
import subprocess
import sys

M = {"HFMA2": "H", "HMUL2": "H", "HADD2": "H", "FFMA": "F", "FMUL": "F", "FADD": "F", "F2FP": "a", "HMNMX2": "a", "PRMT": "a", "SHF": "a",
     "IADD3": "a", "LOP3": "a", "MOV": "a", "MUFU": "X", "LDTM": "T", "ST": "S", "STG": "S", "NOP": ".", "BRA": "B"}


def main():
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
    lines = []
    for l in out.splitlines():
        t = l.strip().split()
        if len(t) > 1 and t[0].startswith("/*") and len(t[0]) == 8:
            t = t[1:]
            op = t[1] if t[0].startswith("@") else t[0]
            lines.append(op.split(".")[0].rstrip(";"))
    first = next(i for i, o in enumerate(lines) if o == "LDTM")
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 380
    s = "".join(M.get(o, "u") for o in lines[first - 2:first - 2 + n])
    for i in range(0, len(s), 100):
        print(s[i:i + 100])
    # burstiness: longest runs
    import itertools
    runs = sorted(((len(list(g)), k) for k, g in itertools.groupby(s)), reverse=True)[:6]
    print("longest runs:", runs, " H:", s.count("H"), " a:", s.count("a"), " total:", len(s))


if __name__ == "__main__":
    main()

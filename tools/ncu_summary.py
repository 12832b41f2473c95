"""Compact per-launch table from an ncu report:  python tools/ncu_summary.py report.ncu-rep > profiles/x.md
(duration, DRAM read/write bytes, DRAM %, tensor-pipe %, issue %, registers, grid/block)."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(k for _, k in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for n, r in enumerate(rows[2:]):
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in KEYS:
            v, u = r[ix[key]], units[ix[key]]
            try:
                f = float(v)
                if u == "Gbyte":
                    cells.append(f"{f * 1e3:.1f} MB")
                elif u == "Mbyte":
                    cells.append(f"{f:.1f} MB")
                elif u == "Kbyte":
                    cells.append(f"{f / 1e3:.2f} MB")
                elif u == "ms":
                    cells.append(f"{f * 1e3:.1f}")
                else:
                    cells.append(f"{f:.1f}" if "." in v else v)
            except ValueError:
                cells.append(v)
        print(f"| {n} | {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()

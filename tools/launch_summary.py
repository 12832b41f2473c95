"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid, block) the number of launches per
forward, the mean duration and the share of the summed kernel time.   python tools/launch_summary.py launches.csv passes > summary.csv"""
import csv
import sys
from collections import defaultdict


def main():
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    k, v, g, b = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[(r[k].replace("void pir::", "").replace("pir::", "").split("(CUtensorMap")[0][:70], r[g], r[b])].append(float(r[v].replace(",", "")))
    tot = sum(sum(x) for x in agg.values())
    print("kernel,grid,block,launches_per_forward,avg_us,share_pct")
    for key, x in sorted(agg.items(), key=lambda t: -sum(t[1])):
        print(f'"{key[0]}","{key[1]}","{key[2]}",{len(x) // passes},{sum(x) / len(x) / 1000:.1f},{100 * sum(x) / tot:.2f}')


if __name__ == "__main__":
    main()

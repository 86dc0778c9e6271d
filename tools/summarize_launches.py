#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share).
   python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        a = agg.setdefault(r[ki], [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel summary of %s (ncu launch list: cold-cache, serialised -- compare SHARES, not absolutes)" % path)
    print("%-34s %6s %12s %11s %7s  %-16s %s" % ("kernel", "n", "total_us", "mean_us", "share", "grid", "block"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        name = k.split("::")[-1].split("(")[0]
        print("%-34s %6d %12.1f %11.1f %6.1f%%  %-16s %s" % (name, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, a[2], a[3]))
    print("total_us %.1f" % tot)


if __name__ == "__main__":
    main(sys.argv[1])

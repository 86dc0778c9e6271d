#!/usr/bin/env python
"""Compact per-kernel table from `ncu -i X.ncu-rep --page raw --csv` output.
   python tools/ncu_table.py raw.csv > profiles/rNN_ncu_summary.txt"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us", 1e-3, "ns"),
    ("dram__bytes_read.sum", "dram_rd_MB", None, None),
    ("dram__bytes_write.sum", "dram_wr_MB", None, None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1, None),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wf%", 1, None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1, None),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%", 1, None),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 1, None),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1, None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1, None),
    ("launch__registers_per_thread", "regs", 1, None),
    ("smsp__inst_executed.sum", "Minst", 1e-6, None),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long", 1, None),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", 1, None),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", 1, None),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar", 1, None),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short", 1, None),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", 1, None),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1) / 1e6


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-22s" % "kernel" + "".join("%10s" % c[1] for c in COLS) + "  grid")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("::")[-1].split("(")[0][:22]
        out = "%-22s" % name
        for key, label, scale, _ in COLS:
            if key not in idx or r[idx[key]] == "":
                out += "%10s" % "-"
                continue
            if label.startswith("dram_"):
                out += "%10.1f" % to_bytes(r[idx[key]], units[idx[key]])
            else:
                v = float(r[idx[key]].replace(",", ""))
                if label == "time_us":
                    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[idx[key]], 1e-3)
                elif scale:
                    v *= scale
                out += "%10.1f" % v
        print(out + "  " + r[idx["Grid Size"]])


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python
"""DRAM traffic per stage from an `ncu --set full` capture of tools/profile_case.py (one chunk of 32 frames):
   ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv;  python tools/ncu_traffic.py raw.csv 32 > profiles/ncu_traffic.json
bench.py reads the result for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernels)."""
import csv
import json
import sys

STAGE_OF = [("k_descriptor", "descriptor"), ("k_support_match", "support_match"), ("k_dcan_border", "support_match"),
            ("k_support_filter", "support_filter"), ("k_incon_first_sweep", "support_filter"), ("k_delaunay_order", "delaunay_device"),
            ("k_delaunay_levels", "delaunay_device"), ("k_post_fused", "post_fused"), ("PostFusedArgs", "post_fused"), ("k_planes", "planes"), ("k_grid", "grid"), ("k_raster", "raster"),
            ("k_dense", "dense_match"), ("DenseArgs", "dense_match"), ("k_lr_check", "lr_check"), ("k_ccl", "remove_small_segments"),
            ("k_gap", "gap_interpolation"), ("k_mean", "adaptive_mean"), ("k_median", "median"), ("k_reproject", "reproject")]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


rows = list(csv.reader(open(sys.argv[1])))
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 32
head, units = rows[0], rows[1]
ki, ri, wi, ti = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"), head.index("gpu__time_duration.sum")
ai = head.index("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
ii = head.index("smsp__issue_active.avg.pct_of_peak_sustained_active")
out = {}
for r in rows[2:]:
    stage = next((s for k, s in STAGE_OF if k in r[ki]), None)
    if stage is None:
        continue
    e = out.setdefault(stage, {"dram_bytes_per_launch": 0.0, "kernels": 0, "ncu_time_us": 0.0})
    e["dram_bytes_per_launch"] += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
    e["kernels"] += 1
    t_us = float(r[ti].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ti], 1e-3)
    e["ncu_time_us"] += t_us
    if t_us > e.get("_longest", 0.0):  # pipe utilisation of the stage's longest kernel
        e["_longest"] = t_us
        e["alu_pipe_pct"] = float(r[ai].replace(",", ""))
        e["issue_active_pct"] = float(r[ii].replace(",", ""))
for e in out.values():
    e.pop("_longest", None)
print(json.dumps({"source": sys.argv[1], "frames_per_launch": frames, "stages": out}, indent=1))

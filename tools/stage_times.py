#!/usr/bin/env python
"""Per-stage CUDA-event times of ONE frame through svb_process at a given size (diagnostic).
   python tools/stage_times.py 3840 2160 511"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

svb = load_package().binding
W, H, dm = (int(x) for x in sys.argv[1:4])
L, R = svb.synth_pair(9, W, H, 0)
p = svb.default_params(svb.MIDDLEBURY, disp_max=dm)
ctx = svb.Context(p, W, H)
ctx.set_stage_timing(True)
for _ in range(3):
    ctx.process(L, R)
st = ctx.stats()
print({k: round(v, 3) for k, v in st["stage_ms"].items() if v > 0}, "delaunay_ms", round(st["delaunay_ms_wall"], 2), "support", st["support_points"])

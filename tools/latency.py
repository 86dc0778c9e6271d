#!/usr/bin/env python
"""Single-frame latency of the drop-in path (what generatePointCloud does per call): BGRA pair in host memory ->
u8 disparity + double3 point cloud in host memory, one frame at a time, synchronously.
   python tools/latency.py [width height]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

svb = load_package().binding
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1242, 375)
frames = [svb.synth_pair(i, W, H, i & 1) for i in range(8)]
bgra = [tuple(np.ascontiguousarray(np.stack([g, g, g, np.full_like(g, 255)], -1)) for g in p) for p in frames]
ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=1)
ctx.set_calibration(np.array([[1, 0, 0, -W / 2.0], [0, 1, 0, -H / 2.0], [0, 0, 0, 0.58 * W], [0, 0, 1.8616, 0]]))
# library-owned pinned result buffers, like generatePointCloud's (dropin.cpp); the inputs stay ordinary (pageable) memory
import ctypes as C

pts = svb.PinnedArray((H * W, 3), np.float64)
dmap = svb.PinnedArray((H, W), np.uint8)
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def call(l, r):
    rc = ctx.lib.svb_point_cloud_bgra(ctx.h, vp(l), vp(r), vp(pts.array), vp(dmap.array), None, None)
    assert rc == 0, rc


for l, r in bgra[:3]:
    call(l, r)
ts = []
for k in range(40):
    l, r = bgra[k % len(bgra)]
    t0 = time.perf_counter()
    call(l, r)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(json.dumps({"size": [W, H], "path": "svb_point_cloud_bgra (BGRA host -> u8 map + double3 cloud host), one frame per call",
                  "ms_median": float(np.median(ts)), "ms_min": float(ts.min()), "fps_median": float(1e3 / np.median(ts))}))
ctx.close()

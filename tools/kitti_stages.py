#!/usr/bin/env python
"""Per-stage CUDA-event times on the reference's own frames (datasets/kitti_mini, tests/golden/kitti_gray.npz tiled to a batch),
single stream so that each stage is timed alone, plus the multi-lane throughput:  python tools/kitti_stages.py [frames] [chunk]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

svb = load_package().binding
n = int(sys.argv[1]) if len(sys.argv) > 1 else 504
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
z = np.load(os.path.join(ROOT, "tests", "golden", "kitti_gray.npz"))
npairs = len([k for k in z.files if k.startswith("L")])
L = np.ascontiguousarray(np.stack([z["L%d" % (i % npairs)] for i in range(n)]))
R = np.ascontiguousarray(np.stack([z["R%d" % (i % npairs)] for i in range(n)]))
H, W = L.shape[1:]
ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=chunk)
ctx.set_calibration(np.array([[1, 0, 0, -738.8], [0, 1, 0, -254.76], [0, 0, 0, 1027.86], [0, 0, 1.8616, 0]]))
ctx.set_stage_timing(True)
ctx.batch_upload(L, R)
flags = svb.OUT_DISPARITY | svb.OUT_POINTS
out = {}
for single in (False, True):
    ctx.set_single_stream(single)
    ctx.batch_run(n, flags)
    ms = 0.0
    for _ in range(2):
        ctx.batch_run(n, flags)
        st = ctx.stats()
        ms += st["gpu_ms_total"]
    out["single_stream" if single else "multi_lane"] = {
        "frames_per_s": round(2 * n / (ms * 1e-3)), "stage_us_per_frame": {k: round(1e3 * v / n, 2) for k, v in st["stage_ms"].items() if v > 0.0005 * n},
        "lists_device": st["delaunay_lists_device"], "lists_host": st["delaunay_lists_host"], "host_ms_per_frame": round(st["delaunay_ms_total"] / n, 4)}
print(json.dumps(out))
ctx.close()

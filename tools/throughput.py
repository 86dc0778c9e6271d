#!/usr/bin/env python
"""Frame-batch throughput at an arbitrary frame size (the parity-test configurations of BASELINE.json other than the
bench workload): python tools/throughput.py 1920 1080 255 [frames] [chunk]
Synthetic pairs, pipeline preset, disparity + point cloud, inputs resident; prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

svb = load_package().binding
W, H, dm = (int(x) for x in sys.argv[1:4])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 128
chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 8
pairs = [svb.synth_pair(i, W, H, i & 1) for i in range(min(n, 16))]
Ls = np.stack([pairs[i % len(pairs)][0] for i in range(n)])
Rs = np.stack([pairs[i % len(pairs)][1] for i in range(n)])
p = svb.default_params(svb.PIPELINE, disp_max=dm)
ctx = svb.Context(p, W, H, chunk=chunk)
ctx.set_calibration(np.array([[1, 0, 0, -W / 2.0], [0, 1, 0, -H / 2.0], [0, 0, 0, 0.58 * W], [0, 0, 1.8616, 0]]))
ctx.set_stage_timing(True)
if os.environ.get("SVB_SINGLE_STREAM") == "1":  # every lane on one stream: the stage times are then exact (no overlap between kernels)
    ctx.set_single_stream(True)
ctx.batch_upload(Ls, Rs)
flags = svb.OUT_DISPARITY | svb.OUT_POINTS
ctx.batch_run(n, flags)
ms = 0.0
steps = 3
for _ in range(steps):
    ctx.batch_run(n, flags)
    st = ctx.stats()
    ms += st["gpu_ms_total"]
print(json.dumps({"size": [W, H], "disp_max": dm, "frames": n, "chunk": chunk, "frames_per_s": n * steps / (ms * 1e-3),
                  "ms_per_frame": ms / (n * steps), "support_points_per_frame": st["support_points"] / n,
                  "host_delaunay_ms_per_frame_cpu": st["delaunay_ms_total"] / n,
                  "stage_us_per_frame": {k: round(1e3 * v / n, 2) for k, v in st["stage_ms"].items() if v > 0}}))
ctx.close()

#!/usr/bin/env python
"""One chunk of the bench workload (32 synthetic KITTI-shape pairs, pipeline preset, disparity + point cloud), run
`reps` times on one stream -- the program ncu is pointed at:
   ncu --set full --import-source on --clock-control none --launch-skip 24 -c 24 -o gpurun_out/prof python tools/profile_case.py
(24 launches per chunk; the first repetition is the warm-up that is skipped)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

svb = load_package().binding
W, H, n = 1242, 375, 32
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pairs = [svb.synth_pair(i, W, H, i & 1) for i in range(n)]
Ls = np.stack([p[0] for p in pairs])
Rs = np.stack([p[1] for p in pairs])
ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=n)
ctx.set_calibration(np.array([[1, 0, 0, -609.5593], [0, 1, 0, -172.854], [0, 0, 0, 721.5377], [0, 0, 1.8616, 0]]))
ctx.set_single_stream(True)
ctx.batch_upload(Ls, Rs)
for _ in range(reps):
    ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
    print("launches", ctx.stats()["kernel_launches"], "frames", ctx.stats()["frames"])
ctx.close()
